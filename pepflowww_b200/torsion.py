"""Side-chain reconstruction with the reference's names (models_con/torsion.py): the step that follows
FlowModel.sample in the sampling scripts (models_con/sample.py:105-108) - SURVEY.md section 8f rank 3.
full_atom_reconstruction and get_heavyatom_mask run hand-written kernels (pf_full_atom_reconstruction,
csrc/pf_recon.cu); CUDA tensors only, no CPU fallback."""
import torch

from . import constants, ops
from .constants import restype_to_heavyatom_masks, torsions_mask  # noqa: F401  (torsion.py:122-124, :230-232)


def _require_cuda(name, t):
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensors required (no CPU fallback)")


def _check_types(aa):
    # the reference gathers from 21-row tables (torsion.py:105-111): PAD (21) or anything outside [0, 20] raises
    if aa.numel() and (int(aa.max()) > 20 or int(aa.min()) < 0):
        raise IndexError("full_atom_reconstruction: residue type outside [0, 20]")


def full_atom_reconstruction(R_bb, t_bb, angles, aa):
    """(pos14 [B,N,14,3], R [B,N,6,3,3], t [B,N,6,3]): atom14 coordinates and the backbone / psi / chi1-4 frames from
    backbone frames, torsions in [0, 2 pi) and residue types (models_con/torsion.py:140-226)."""
    _require_cuda("full_atom_reconstruction", R_bb)
    if R_bb.shape[:2] != aa.shape or angles.shape != aa.shape + (5,) or t_bb.shape != aa.shape + (3,):
        raise ValueError("full_atom_reconstruction: shapes must be R [B,N,3,3], t [B,N,3], angles [B,N,5], aa [B,N]")
    _check_types(aa)
    pos14, R, t, _ = ops.full_atom_reconstruction(R_bb, t_bb, angles, aa, constants.rigid_tables(R_bb.device))
    return pos14, R, t


def get_heavyatom_mask(aa):
    """[B,N,15] bool: which atom slots a residue type has, OXT excluded (models_con/torsion.py:126-138)."""
    return restype_to_heavyatom_masks.to(aa.device)[aa.flatten()].reshape(*aa.shape, 15)


def get_torsion_angle(pos14, aa):
    """(torsion [..., 5], torsion_mask [..., 5]): psi (N, CA, C, O) and chi1-4 in [0, 2 pi) from atom14 coordinates
    pos14 [..., A >= 14, 3] and residue types aa [...] (models_con/torsion.py:49-66, there one residue at a time in a
    Python loop); one launch of pf_torsion_angles for the whole batch."""
    _require_cuda("get_torsion_angle", pos14)
    if pos14.shape[:-2] != aa.shape or pos14.shape[-1] != 3 or not 14 <= pos14.shape[-2] <= 15:
        raise ValueError("get_torsion_angle: shapes must be pos14 [..., 14|15, 3], aa [...]")
    return ops.torsion_angles(pos14, aa, constants.rigid_tables(pos14.device))


def reconstruct_side_chains(samples):
    """pos_ha [B,N,15,3] and mask [B,N,15] from the last trajectory entry of FlowModel.sample - what
    save_samples_sc does before writing PDB files (models_con/sample.py:105-108) - in one launch."""
    R, t, ang, aa = samples["rotmats"], samples["trans"], samples["angles"], samples["seqs"]
    _require_cuda("reconstruct_side_chains", R)
    _check_types(aa)
    pos14, _, _, mask = ops.full_atom_reconstruction(R, t, ang, aa, constants.rigid_tables(R.device),
                                                     want_frames=False, want_mask=True)
    return torch.nn.functional.pad(pos14, (0, 0, 0, 1), value=0.0), mask
