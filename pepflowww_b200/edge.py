"""EdgeEmbedder (reference: models_con/edge.py:15-112): once per sample; same parameter names/shapes.
SURVEY section 8(f) rank 1.  On a CUDA device without autograd (sampling) the whole module is ONE fused kernel
(pf_edge_embed, csrc/pf_embed.cu): atom coordinates -> [B,L,L,64] pair features, nothing of size L^2 x 225 in HBM.
With autograd (training) or on CPU tensors the reference formulation below runs as torch ops, the pair dimension
processed in row chunks so the [B,L,L,225] distance features never exceed `chunk_bytes` (the reference
materialises them ~5x, 21 GB transient at B=64, L=271)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .constants import AA, BBHeavyAtom
from .geometry import dihedral_from_four_points
from .layers import AngularEncoding


class EdgeEmbedder(nn.Module):
    def __init__(self, feat_dim, max_num_atoms, max_aa_types=22, max_relpos=32, num_bins=16):
        super().__init__()
        self.max_num_atoms, self.max_aa_types, self.max_relpos, self.num_bins = max_num_atoms, max_aa_types, max_relpos, num_bins
        self.aa_pair_embed = nn.Embedding(max_aa_types * max_aa_types, feat_dim)
        self.relpos_embed = nn.Embedding(2 * max_relpos + 1, feat_dim)
        self.aapair_to_distcoef = nn.Embedding(max_aa_types * max_aa_types, max_num_atoms * max_num_atoms)
        nn.init.zeros_(self.aapair_to_distcoef.weight)
        self.distance_embed = nn.Sequential(nn.Linear(max_num_atoms * max_num_atoms, feat_dim), nn.ReLU(),
                                            nn.Linear(feat_dim, feat_dim), nn.ReLU())
        self.dihedral_embed = AngularEncoding()
        infeat = feat_dim * 3 + self.dihedral_embed.get_out_dim(2)
        self.out_mlp = nn.Sequential(nn.Linear(infeat, feat_dim), nn.ReLU(), nn.Linear(feat_dim, feat_dim), nn.ReLU(),
                                     nn.Linear(feat_dim, feat_dim))
        self.chunk_bytes = 1 << 30

    def _rows(self, sl, aa, res_nb, chain_nb, pos, mask_atoms, mask_residue, structure_mask):
        """Edge features for query rows `sl` against all columns: [N, l, L, F]."""
        N, L = aa.shape
        l = sl.stop - sl.start
        aa_pair = aa[:, sl, None] * self.max_aa_types + aa[:, None, :]
        f_aa = self.aa_pair_embed(aa_pair)
        same = chain_nb[:, sl, None] == chain_nb[:, None, :]
        rel = torch.clamp(res_nb[:, sl, None] - res_nb[:, None, :], min=-self.max_relpos, max=self.max_relpos)
        f_rel = self.relpos_embed(rel + self.max_relpos) * same[..., None]
        d = torch.linalg.norm(pos[:, sl, None, :, None] - pos[:, None, :, None, :], dim=-1, ord=2).reshape(N, l, L, -1) / 10
        c = F.softplus(self.aapair_to_distcoef(aa_pair))
        g = torch.exp(-1 * c * d ** 2)
        m = (mask_atoms[:, sl, None, :, None] * mask_atoms[:, None, :, None, :]).reshape(N, l, L, -1)
        f_d = self.distance_embed(g * m)
        pN, pCA, pC = pos[:, :, BBHeavyAtom.N], pos[:, :, BBHeavyAtom.CA], pos[:, :, BBHeavyAtom.C]
        rows = lambda x: x[:, sl, None].expand(N, l, L, 3)
        cols = lambda x: x[:, None, :].expand(N, l, L, 3)
        phi = dihedral_from_four_points(rows(pC), cols(pN), cols(pCA), cols(pC))
        psi = dihedral_from_four_points(rows(pN), rows(pCA), rows(pC), cols(pN))
        f_h = self.dihedral_embed(torch.stack([phi, psi], dim=-1))
        if structure_mask is not None:
            psm = (structure_mask[:, sl, None] * structure_mask[:, None, :])[..., None]
            f_d, f_h = f_d * psm, f_h * psm
        out = self.out_mlp(torch.cat([f_aa, f_rel, f_d, f_h], dim=-1))
        return out * (mask_residue[:, sl, None] * mask_residue[:, None, :])[..., None]

    def _kernel_constants(self):
        """Constants of pf_edge_embed in prototype order: softplus(coef), the two embedding tables pushed through the
        first output layer, transposed weights, biases."""
        F_ = self.aa_pair_embed.weight.shape[1]
        d0, d2 = self.distance_embed[0], self.distance_embed[2]
        o0, o2, o4 = self.out_mlp[0], self.out_mlp[2], self.out_mlp[4]
        w = o0.weight
        return (F.softplus(self.aapair_to_distcoef.weight), self.aa_pair_embed.weight @ w[:, :F_].t(),
                self.relpos_embed.weight @ w[:, F_:2 * F_].t(), d0.weight.t().contiguous(), d0.bias,
                d2.weight.t().contiguous(), d2.bias, w[:, 2 * F_:3 * F_].t().contiguous(), w[:, 3 * F_:].t().contiguous(),
                o0.bias, o2.weight.t().contiguous(), o2.bias, o4.weight.t().contiguous(), o4.bias)

    def forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
        """The fused kernel (pf_edge_embed).  CUDA tensors, no autograd - there is no silent torch or CPU fallback; the
        differentiable formulation of the training path is forward_autograd()."""
        if not pos_atoms.is_cuda:
            raise RuntimeError("EdgeEmbedder.forward runs the CUDA kernel (no CPU fallback); got a CPU tensor")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("EdgeEmbedder.forward runs the inference kernel; wrap the call in torch.no_grad() "
                               "(the autograd training path is EdgeEmbedder.forward_autograd)")
        if not (self.max_num_atoms == 15 and self.max_aa_types == 22 and self.max_relpos == 32 and
                self.aa_pair_embed.weight.shape[1] == 64 and self.dihedral_embed.num_funcs == 3):
            raise ValueError("pf_edge_embed is specialised to 15 atoms / 22 residue types / relpos 32 / 64 channels")
        from . import ops
        if sequence_mask is not None:
            aa = torch.where(sequence_mask, aa, torch.full_like(aa, fill_value=int(AA.UNK)))
        return ops.edge_embed(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, self._kernel_constants())

    def forward_autograd(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
        """Same contract in differentiable torch ops over the same parameters (models_con/edge.py:39-112): the gradient
        path of FlowModel.forward / train_ddp.py."""
        N, L = aa.size()
        A = self.max_num_atoms
        pos_atoms, mask_atoms = pos_atoms[:, :, :A], mask_atoms[:, :, :A]
        mask_residue = mask_atoms[:, :, BBHeavyAtom.CA]
        if sequence_mask is not None:
            aa = torch.where(sequence_mask, aa, torch.full_like(aa, fill_value=int(AA.UNK)))
        per_row = N * L * A * A * 4 * 6
        step = max(1, min(L, self.chunk_bytes // max(per_row, 1)))
        outs = [self._rows(slice(s, min(L, s + step)), aa, res_nb, chain_nb, pos_atoms, mask_atoms, mask_residue,
                           structure_mask) for s in range(0, L, step)]
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
