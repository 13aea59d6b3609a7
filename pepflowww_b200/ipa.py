"""Blocks of the denoiser with the reference's class names, parameter names and shapes
(models_con/ipa_pytorch.py: Linear :116-181, StructureModuleTransition :184-206, EdgeTransition :209-248,
InvariantPointAttention :251-484, BackboneUpdate :544-572) so checkpoints load unchanged.

The nn.Modules only own parameters; every forward() enqueues hand-written sm_100a kernels through the C ABI
(pepflowww_b200/ops.py).  There is no PyTorch or CPU fallback in these forward paths.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .rigid import Rigid

_TRUNC_STD = 0.8796256610342398  # std of a unit normal truncated to [-2, 2]


def _trunc_normal_(w, scale):
    fan_in = w.shape[1]
    std = math.sqrt(scale / max(1, fan_in)) / _TRUNC_STD
    with torch.no_grad():
        nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2.0 * std, b=2.0 * std)


class Linear(nn.Linear):
    """nn.Linear with the reference's named initialisers ('default' LeCun, 'relu' He, 'final' zeros, ...)."""

    def __init__(self, in_dim, out_dim, bias=True, init="default"):
        super().__init__(in_dim, out_dim, bias=bias)
        with torch.no_grad():
            if bias:
                self.bias.fill_(0)
            if init == "default":
                _trunc_normal_(self.weight, 1.0)
            elif init == "relu":
                _trunc_normal_(self.weight, 2.0)
            elif init == "glorot":
                nn.init.xavier_uniform_(self.weight, gain=1)
            elif init == "gating":
                self.weight.fill_(0.0)
                if bias:
                    self.bias.fill_(1.0)
            elif init == "normal":
                nn.init.kaiming_normal_(self.weight, nonlinearity="linear")
            elif init == "final":
                self.weight.fill_(0.0)
            else:
                raise ValueError("Invalid init string.")

    def forward(self, x):
        return ops.linear(x, self.weight, self.bias)


class StructureModuleTransition(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.c = c
        self.linear_1 = Linear(c, c, init="relu")
        self.linear_2 = Linear(c, c, init="relu")
        self.linear_3 = Linear(c, c, init="final")
        self.relu = nn.ReLU()
        self.ln = nn.LayerNorm(c)

    def forward(self, s):
        h = ops.linear(s, self.linear_1.weight, self.linear_1.bias, act=1)
        h = ops.linear(h, self.linear_2.weight, self.linear_2.bias, act=1)
        h = ops.linear(h, self.linear_3.weight, self.linear_3.bias)
        return ops.add_layernorm(s, h, self.ln.weight, self.ln.bias)


class EdgeTransition(nn.Module):
    def __init__(self, *, node_embed_size, edge_embed_in, edge_embed_out, num_layers=2, node_dilation=2):
        super().__init__()
        bias_embed_size = node_embed_size // node_dilation
        self.initial_embed = Linear(node_embed_size, bias_embed_size, init="relu")
        hidden = bias_embed_size * 2 + edge_embed_in
        layers = []
        for _ in range(num_layers):
            layers.append(Linear(hidden, hidden, init="relu"))
            layers.append(nn.ReLU())
        self.trunk = nn.Sequential(*layers)
        self.final_layer = Linear(hidden, edge_embed_out, init="final")
        self.layer_norm = nn.LayerNorm(edge_embed_out)
        if (node_embed_size, edge_embed_in, edge_embed_out, num_layers, node_dilation) != (128, 64, 64, 2, 2):
            raise ValueError("EdgeTransition kernel is specialised to node 128 / edge 64 / 2 layers / dilation 2")

    def forward(self, node_embed, edge_embed, edge_mask_rows=None, out=None):
        """[B,L,128], [B,L,L,64] -> [B,L,L,64].  `edge_mask_rows` ([B,L]) fuses the `* edge_mask` of ga.py:118."""
        B, L, _ = node_embed.shape
        mask = edge_mask_rows if edge_mask_rows is not None else torch.ones(B, L, device=node_embed.device)
        return ops.edge_transition(node_embed, edge_embed, self.initial_embed.weight, self.initial_embed.bias,
                                   self.trunk[0].weight, self.trunk[0].bias, self.trunk[2].weight, self.trunk[2].bias,
                                   self.final_layer.weight, self.final_layer.bias, self.layer_norm.weight,
                                   self.layer_norm.bias, mask.float(), out=out)


class InvariantPointAttention(nn.Module):
    def __init__(self, ipa_conf, inf=1e5, eps=1e-8):
        super().__init__()
        self._ipa_conf = ipa_conf
        self.c_s, self.c_z, self.c_hidden = ipa_conf.c_s, ipa_conf.c_z, ipa_conf.c_hidden
        self.no_heads, self.no_qk_points, self.no_v_points = ipa_conf.no_heads, ipa_conf.no_qk_points, ipa_conf.no_v_points
        if (self.c_s, self.c_z, self.c_hidden, self.no_heads, self.no_qk_points, self.no_v_points) != (128, 64, 128, 8, 8, 12):
            raise ValueError("IPA kernels are specialised to c_s 128, c_z 64, c_hidden 128, 8 heads, 8/12 points")
        if inf != 1e5 or eps != 1e-8:
            raise ValueError("IPA kernels are specialised to inf=1e5, eps=1e-8")
        self.inf, self.eps = inf, eps
        hc = self.c_hidden * self.no_heads
        self.linear_q = Linear(self.c_s, hc)
        self.linear_kv = Linear(self.c_s, 2 * hc)
        self.linear_q_points = Linear(self.c_s, self.no_heads * self.no_qk_points * 3)
        self.linear_kv_points = Linear(self.c_s, self.no_heads * (self.no_qk_points + self.no_v_points) * 3)
        self.linear_b = Linear(self.c_z, self.no_heads)
        self.down_z = Linear(self.c_z, self.c_z // 4)
        self.head_weights = nn.Parameter(torch.zeros(self.no_heads))
        with torch.no_grad():
            self.head_weights.fill_(0.541324854612918)  # softplus^-1(1)
        concat_out_dim = self.c_z // 4 + self.c_hidden + self.no_v_points * 4
        self.linear_out = Linear(self.no_heads * concat_out_dim, self.c_s, init="final")

    def packed_projection(self):
        """cat(linear_q, linear_kv, linear_q_points, linear_kv_points) -> W [3744,128], b [3744]."""
        w = torch.cat([self.linear_q.weight, self.linear_kv.weight, self.linear_q_points.weight,
                       self.linear_kv_points.weight], dim=0).contiguous()
        b = torch.cat([self.linear_q.bias, self.linear_kv.bias, self.linear_q_points.bias,
                       self.linear_kv_points.bias], dim=0).contiguous()
        return w, b

    def scaled_head_weights(self):
        return (torch.nn.functional.softplus(self.head_weights) *
                math.sqrt(1.0 / (3 * (self.no_qk_points * 9.0 / 2)))).contiguous()

    def forward(self, s, z, r: Rigid, mask):
        """s [B,L,128], z [B,L,L,64], r Rigid over [B,L], mask [B,L] -> [B,L,128]."""
        w, b = self.packed_projection()
        rot = getattr(r.get_rots(), "_cached_rot", None)
        if rot is None:
            rot = r.get_rots().get_rot_mats()
        trans = r.get_trans()
        m = mask.float()
        proj = ops.linear(s, w, b)
        pts = ops.ipa_points(proj, rot, trans)
        feats = ops.ipa_attention(proj, pts, z, self.linear_b.weight, self.linear_b.bias, self.down_z.weight,
                                  self.down_z.bias, self.scaled_head_weights(), rot, trans, m)
        return ops.linear(feats, self.linear_out.weight, self.linear_out.bias)


class BackboneUpdate(nn.Module):
    def __init__(self, c_s, use_rot_updates):
        super().__init__()
        self.c_s = c_s
        self._use_rot_updates = use_rot_updates
        self.linear = Linear(c_s, 6 if use_rot_updates else 3, init="final")

    def forward(self, s):
        return ops.linear(s, self.linear.weight, self.linear.bias)
