"""NodeEmbedder (reference: models_con/node.py:11-105): once per sample; same parameter names/shapes.
SURVEY section 8(f) rank 2.  On a CUDA device without autograd (sampling) the module is ONE fused kernel
(pf_node_embed, csrc/pf_embed.cu); with autograd or on CPU tensors the reference formulation below runs as torch ops."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .constants import AA, BBHeavyAtom
from .geometry import construct_3d_basis, get_backbone_dihedral_angles, global_to_local
from .layers import AngularEncoding


class NodeEmbedder(nn.Module):
    def __init__(self, feat_dim, max_num_atoms, max_aa_types=22):
        super().__init__()
        self.max_num_atoms, self.max_aa_types, self.feat_dim = max_num_atoms, max_aa_types, feat_dim
        self.aatype_embed = nn.Embedding(max_aa_types, feat_dim)
        self.dihed_embed = AngularEncoding()
        infeat = feat_dim + max_aa_types * max_num_atoms * 3 + self.dihed_embed.get_out_dim(3)
        self.mlp = nn.Sequential(nn.Linear(infeat, feat_dim * 2), nn.ReLU(), nn.Linear(feat_dim * 2, feat_dim), nn.ReLU(),
                                 nn.Linear(feat_dim, feat_dim), nn.ReLU(), nn.Linear(feat_dim, feat_dim))

    def _kernel_constants(self):
        """Constants of pf_node_embed in prototype order (the embedding lookup pushed through the first layer)."""
        F_ = self.feat_dim
        l0, l2, l4, l6 = self.mlp[0], self.mlp[2], self.mlp[4], self.mlp[6]
        w = l0.weight
        nc = self.max_aa_types * self.max_num_atoms * 3
        return (self.aatype_embed.weight @ w[:, :F_].t() + l0.bias, w[:, F_:F_ + nc].t().contiguous(),
                w[:, F_ + nc:].t().contiguous(), l2.weight.t().contiguous(), l2.bias, l4.weight.t().contiguous(), l4.bias,
                l6.weight.t().contiguous(), l6.bias)

    def forward(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
        """The fused kernel (pf_node_embed).  CUDA tensors, no autograd - no silent torch or CPU fallback; the
        differentiable formulation of the training path is forward_autograd()."""
        if not pos_atoms.is_cuda:
            raise RuntimeError("NodeEmbedder.forward runs the CUDA kernel (no CPU fallback); got a CPU tensor")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("NodeEmbedder.forward runs the inference kernel; wrap the call in torch.no_grad() "
                               "(the autograd training path is NodeEmbedder.forward_autograd)")
        if not (self.max_num_atoms == 15 and self.max_aa_types == 22 and self.feat_dim == 128 and
                self.dihed_embed.num_funcs == 3 and pos_atoms.shape[2] >= 15):
            raise ValueError("pf_node_embed is specialised to 15 atoms / 22 residue types / 128 channels")
        from . import ops
        if sequence_mask is not None:
            aa = torch.where(sequence_mask, aa, torch.full_like(aa, fill_value=int(AA.UNK)))
        return ops.node_embed(aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask, self._kernel_constants())

    def forward_autograd(self, aa, res_nb, chain_nb, pos_atoms, mask_atoms, structure_mask=None, sequence_mask=None):
        """Same contract in differentiable torch ops over the same parameters (models_con/node.py:35-105): the gradient
        path of FlowModel.forward / train_ddp.py."""
        N, L = aa.size()
        A = self.max_num_atoms
        mask_residue = mask_atoms[:, :, BBHeavyAtom.CA]
        pos_atoms, mask_atoms = pos_atoms[:, :, :A], mask_atoms[:, :, :A]
        if sequence_mask is not None:
            aa = torch.where(sequence_mask, aa, torch.full_like(aa, fill_value=int(AA.UNK)))
        aa_feat = self.aatype_embed(aa)
        R = construct_3d_basis(pos_atoms[:, :, BBHeavyAtom.CA], pos_atoms[:, :, BBHeavyAtom.C], pos_atoms[:, :, BBHeavyAtom.N])
        crd = global_to_local(R, pos_atoms[:, :, BBHeavyAtom.CA], pos_atoms)
        crd = torch.where(mask_atoms[..., None], crd, torch.zeros_like(crd))
        # the local coordinates are placed in the slot of this residue's amino-acid type (22 x A x 3 wide)
        crd_feat = crd.new_zeros(N, L, self.max_aa_types, A * 3)
        crd_feat.scatter_(2, aa[:, :, None, None].expand(N, L, 1, A * 3), crd.reshape(N, L, 1, A * 3))
        crd_feat = crd_feat.reshape(N, L, self.max_aa_types * A * 3)
        if structure_mask is not None:
            crd_feat = crd_feat * structure_mask[:, :, None]
        dih, dih_mask = get_backbone_dihedral_angles(pos_atoms, chain_nb=chain_nb, res_nb=res_nb, mask=mask_residue)
        dih_feat = (self.dihed_embed(dih[:, :, :, None]) * dih_mask[:, :, :, None]).reshape(N, L, -1)
        if structure_mask is not None:
            dm = structure_mask & torch.roll(structure_mask, shifts=+1, dims=1) & torch.roll(structure_mask, shifts=-1, dims=1)
            dih_feat = dih_feat * dm[:, :, None]
        out = self.mlp(torch.cat([aa_feat, crd_feat, dih_feat], dim=-1))
        return out * mask_residue[:, :, None]
