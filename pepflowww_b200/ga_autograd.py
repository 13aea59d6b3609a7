"""Differentiable formulation of the denoiser for the TRAINING path (train_ddp.py:94-168 calls FlowModel.forward
with gradients).  The sampling path never comes here: GAEncoder.forward runs the
hand-written kernels and refuses to run with gradients enabled.  Hand-written backward kernels for the IPA core and
the edge transition are the remaining part of that row; until then the gradient comes from autograd over this
torch-op restatement of models_con/ga.py:87-127, which reads the same nn.Parameters the kernels read - so a checkpoint
trained through it samples through the kernels unchanged.  Works on CUDA and (for the gloo tests) CPU tensors."""
import math

import torch
import torch.nn.functional as F

from .utils_time import get_time_embedding


def _quat_to_rot(q):
    # openfold/utils/rigid_utils.py:185-205
    a, b, c, d = q.unbind(-1)
    return torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
                        2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
                        2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d],
                       dim=-1).reshape(q.shape[:-1] + (3, 3))


@torch.no_grad()
def _rot_to_quat(R):
    # openfold/utils/rigid_utils.py:208-227; the noisy input frames carry no gradient
    m = lambda i, j: R[..., i, j]
    K = torch.stack([m(0, 0) + m(1, 1) + m(2, 2), m(2, 1) - m(1, 2), m(0, 2) - m(2, 0), m(1, 0) - m(0, 1),
                     m(2, 1) - m(1, 2), m(0, 0) - m(1, 1) - m(2, 2), m(0, 1) + m(1, 0), m(0, 2) + m(2, 0),
                     m(0, 2) - m(2, 0), m(0, 1) + m(1, 0), m(1, 1) - m(0, 0) - m(2, 2), m(1, 2) + m(2, 1),
                     m(1, 0) - m(0, 1), m(0, 2) + m(2, 0), m(1, 2) + m(2, 1), m(2, 2) - m(0, 0) - m(1, 1)],
                    dim=-1).reshape(R.shape[:-2] + (4, 4)) / 3.0
    if R.is_cuda:
        # the kernel's rot -> quat (power iteration on the same K matrix, pf_rigid_update with a zero update): no cuSOLVER
        # call, no host synchronisation - the training iteration stays capturable in a CUDA graph
        from . import ops
        z3 = torch.zeros(*R.shape[:-2], 3, device=R.device)
        z6 = torch.zeros(*R.shape[:-2], 6, device=R.device)
        return ops.rigid_update(None, R.contiguous(), z3, z6, torch.ones(R.shape[:-2], device=R.device))[0]
    return torch.linalg.eigh(K)[1][..., -1]


def _lin(layer, x):
    # the package's Linear.forward launches the inference kernel; the autograd path goes through F.linear
    return F.linear(x, layer.weight, layer.bias)


def _ipa(mod, s, z, R, t, mask):
    """InvariantPointAttention.forward, models_con/ipa_pytorch.py:316-484, on the module's parameters."""
    B, L, _ = s.shape
    H, C, PQ, PV = mod.no_heads, mod.c_hidden, mod.no_qk_points, mod.no_v_points
    q = _lin(mod.linear_q, s).view(B, L, H, C)
    k, v = _lin(mod.linear_kv, s).view(B, L, H, 2 * C).split(C, dim=-1)

    def global_points(lin, n):   # x | y | z planes, head-major inside a plane (:360-387), then Rigid.apply
        loc = _lin(lin, s).view(B, L, 3, H * n).transpose(-1, -2)
        return (loc @ R.transpose(-1, -2) + t[:, :, None, :]).view(B, L, H, n, 3)

    q_pts = global_points(mod.linear_q_points, PQ)
    k_pts, v_pts = global_points(mod.linear_kv_points, PQ + PV).split([PQ, PV], dim=3)
    logits = torch.einsum("bihc,bjhc->bhij", q, k) * math.sqrt(1.0 / (3 * C))
    logits = logits + math.sqrt(1.0 / 3) * _lin(mod.linear_b, z).permute(0, 3, 1, 2)
    d2 = (q_pts[:, :, None] - k_pts[:, None, :]).square().sum(-1)                       # [B,L,L,H,PQ]
    logits = logits - 0.5 * (d2 * mod.scaled_head_weights()[:, None]).sum(-1).permute(0, 3, 1, 2)
    logits = logits + (mod.inf * (mask[:, :, None] * mask[:, None, :] - 1))[:, None]
    a = torch.softmax(logits, dim=-1)
    o = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, L, H * C)
    o_pt = torch.einsum("bhij,bjhpx->bihpx", a, v_pts) - t[:, :, None, None, :]
    o_pt = torch.einsum("blji,blhpj->blhpi", R, o_pt).reshape(B, L, H * PV, 3)         # Rigid.invert_apply (:455)
    o_norm = torch.sqrt(o_pt.square().sum(-1) + mod.eps)
    o_pair = torch.einsum("bhij,bijc->bihc", a, _lin(mod.down_z, z)).reshape(B, L, -1)
    return _lin(mod.linear_out, torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_norm, o_pair], dim=-1))


def _mlp(seq, x):
    # nn.Sequential of nn.Linear / nn.ReLU; F.linear keeps it independent of Module.forward overrides
    for layer in seq:
        x = _lin(layer, x) if isinstance(layer, torch.nn.Linear) else layer(x)
    return x


def denoiser_autograd(enc, t, rotmats_t, trans_t, angles_t, seqs_t, node_embed, edge_embed, generate_mask, res_mask):
    """GAEncoder.forward (models_con/ga.py:87-127) as differentiable torch ops over `enc`'s parameters.
    Returns (rotmats [B,L,3,3], trans [B,L,3], angles [B,L,5] in [0, 2 pi), logits [B,L,20])."""
    B, L = seqs_t.shape
    m = res_mask.to(node_embed.dtype)
    em = m[:, None, :] * m[:, :, None]
    temb = get_time_embedding(t[:, 0], enc.feat_dim, max_positions=2056)[:, None, :].expand(B, L, -1)
    x = torch.cat([node_embed, enc.current_seq_embedder(seqs_t), temb, enc.angles_embedder(angles_t)], dim=-1)
    s = _mlp(enc.res_feat_mixer, x) * m[..., None]
    R, tr, quat = rotmats_t.float(), trans_t, None
    z = edge_embed
    pad = m <= 0
    tk = enc.trunk
    nb = enc._ipa_conf.num_blocks
    for b in range(nb):
        s = tk[f"ipa_ln_{b}"](s + _ipa(tk[f"ipa_{b}"], s, z, R, tr, m) * m[..., None])
        y = tk[f"seq_tfmr_{b}"](s, src_key_padding_mask=pad)
        s = s + F.linear(y, tk[f"post_tfmr_{b}"].weight, tk[f"post_tfmr_{b}"].bias)
        s = _node_transition(tk[f"node_transition_{b}"], s) * m[..., None]
        lin = tk[f"bb_update_{b}"].linear
        upd = F.linear(s * m[..., None], lin.weight, lin.bias)
        # Rigid.compose_q_update_vec (openfold/utils/rigid_utils.py:1039-1063, :587-616, normalise :331-332)
        if quat is None:
            quat = _rot_to_quat(R)
        qa, qb, qc, qd = quat.unbind(-1)
        ux, uy, uz = (upd[..., :3] * m[..., None]).unbind(-1)
        quat = quat + torch.stack([-qb * ux - qc * uy - qd * uz, qa * ux + qc * uz - qd * uy,
                                   qa * uy - qb * uz + qd * ux, qa * uz + qb * uy - qc * ux], dim=-1)
        quat = quat / torch.linalg.norm(quat, dim=-1, keepdim=True)
        tr = tr + torch.einsum("blij,blj->bli", R, upd[..., 3:]) * m[..., None]      # rotated by the OLD frame
        R = _quat_to_rot(quat)
        if b < nb - 1:
            z = _edge_transition(tk[f"edge_transition_{b}"], s, z) * em[..., None]
    logits = _mlp(enc.seq_net, s)
    angles = torch.remainder(_mlp(enc.angle_net, s), 2 * math.pi)
    return R, tr, angles, logits


def _node_transition(mod, s):
    h = F.relu(_lin(mod.linear_1, s))
    h = F.relu(_lin(mod.linear_2, h))
    return F.layer_norm(s + _lin(mod.linear_3, h), (s.shape[-1],), mod.ln.weight, mod.ln.bias, mod.ln.eps)


def _edge_transition(mod, s, z):
    B, L, _ = s.shape
    e = _lin(mod.initial_embed, s)
    x = torch.cat([z, e[:, :, None, :].expand(B, L, L, -1), e[:, None, :, :].expand(B, L, L, -1)], dim=-1)
    h = F.relu(_lin(mod.trunk[2], F.relu(_lin(mod.trunk[0], x))))
    y = _lin(mod.final_layer, h + x)
    return F.layer_norm(y, (y.shape[-1],), mod.layer_norm.weight, mod.layer_norm.bias, mod.layer_norm.eps)
