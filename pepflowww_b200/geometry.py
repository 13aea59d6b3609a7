"""Geometry helpers used once per sample by encode() and the embedders (reference:
pepflow/modules/common/geometry.py: construct_3d_basis :89-111, global_to_local :136-155,
dihedral_from_four_points :296-313, get_backbone_dihedral_angles :355-390, pairwise_dihedrals :393-418;
pepflow/modules/common/topology.py:5-24).  Plain torch ops on whatever device the batch lives on; these are
the SURVEY section 8(f) "next" rows, not the per-step hot path."""
import torch
import torch.nn.functional as F

from .constants import BBHeavyAtom


def normalize_vector(v, dim, eps=1e-6):
    return v / (torch.linalg.norm(v, ord=2, dim=dim, keepdim=True) + eps)


def construct_3d_basis(center, p1, p2):
    e1 = normalize_vector(p1 - center, dim=-1)
    v2 = p2 - center
    e2 = normalize_vector(v2 - (e1 * v2).sum(dim=-1, keepdim=True) * e1, dim=-1)
    e3 = torch.cross(e1, e2, dim=-1)
    return torch.stack([e1, e2, e3], dim=-1)


def global_to_local(R, t, q):
    """p = R^T (q - t) for q [N, L, ..., 3]."""
    shape = q.shape
    q = q.reshape(shape[0], shape[1], -1, 3)
    p = torch.einsum("nlji,nlaj->nlai", R, q - t[:, :, None, :])
    return p.reshape(shape)


def dihedral_from_four_points(p0, p1, p2, p3):
    v0, v1, v2 = p2 - p1, p0 - p1, p3 - p2
    u1 = torch.cross(v0, v1, dim=-1)
    n1 = u1 / torch.linalg.norm(u1, dim=-1, keepdim=True)
    u2 = torch.cross(v0, v2, dim=-1)
    n2 = u2 / torch.linalg.norm(u2, dim=-1, keepdim=True)
    sgn = torch.sign((torch.cross(v1, v2, dim=-1) * v0).sum(-1))
    return torch.nan_to_num(sgn * torch.acos((n1 * n2).sum(-1).clamp(min=-0.999999, max=0.999999)))


def get_terminus_flag(chain_nb, res_nb, mask):
    consec = ((res_nb[:, 1:] - res_nb[:, :-1]).abs() == 1) & (chain_nb[:, 1:] == chain_nb[:, :-1]) & mask[:, :-1]
    n_term = F.pad(~consec, pad=(1, 0), value=True)
    c_term = F.pad(~consec, pad=(0, 1), value=True)
    return n_term, c_term


def get_backbone_dihedral_angles(pos_atoms, chain_nb, res_nb, mask):
    pN, pCA, pC = pos_atoms[:, :, BBHeavyAtom.N], pos_atoms[:, :, BBHeavyAtom.CA], pos_atoms[:, :, BBHeavyAtom.C]
    n_term, c_term = get_terminus_flag(chain_nb, res_nb, mask)
    omega = F.pad(dihedral_from_four_points(pCA[:, :-1], pC[:, :-1], pN[:, 1:], pCA[:, 1:]), pad=(1, 0), value=0)
    phi = F.pad(dihedral_from_four_points(pC[:, :-1], pN[:, 1:], pCA[:, 1:], pC[:, 1:]), pad=(1, 0), value=0)
    psi = F.pad(dihedral_from_four_points(pN[:, :-1], pCA[:, :-1], pC[:, :-1], pN[:, 1:]), pad=(0, 1), value=0)
    m = torch.stack([~n_term, ~n_term, ~c_term], dim=-1)
    return torch.stack([omega, phi, psi], dim=-1) * m, m


def pairwise_dihedrals(pos_atoms):
    N, L = pos_atoms.shape[:2]
    pN, pCA, pC = pos_atoms[:, :, BBHeavyAtom.N], pos_atoms[:, :, BBHeavyAtom.CA], pos_atoms[:, :, BBHeavyAtom.C]
    rows = lambda x: x[:, :, None].expand(N, L, L, 3)
    cols = lambda x: x[:, None, :].expand(N, L, L, 3)
    phi = dihedral_from_four_points(rows(pC), cols(pN), cols(pCA), cols(pC))
    psi = dihedral_from_four_points(rows(pN), rows(pCA), rows(pC), cols(pN))
    return torch.stack([phi, psi], dim=-1)


def reconstruct_backbone(R, t, aa, chain_nb, res_nb, mask):
    """N, CA, C, O [N, L, 4, 3] from backbone frames and residue types (reference geometry.py:446-489; the step after
    FlowModel.sample in models_con/sample.py:46,77).  One kernel (pf_reconstruct_backbone); CUDA tensors only."""
    from . import constants, ops
    if not R.is_cuda:
        raise RuntimeError("reconstruct_backbone: CUDA tensors required (no CPU fallback)")
    return ops.reconstruct_backbone(R, t, aa, chain_nb, res_nb, mask, constants.rigid_tables(R.device))
