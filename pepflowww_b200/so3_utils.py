"""SO(3) maps with the reference's names (data/so3_utils.py:143-164 rotvec_to_rotmat, :167-254
rotmat_to_rotvec, :486-497 calc_rot_vf, :500-520 geodesic_t), each one CUDA kernel."""
import torch

from . import ops


def rotvec_to_rotmat(rotation_vectors, tol=1e-7):
    return ops.so3_exp(rotation_vectors)


def rotmat_to_rotvec(rotation_matrices):
    return ops.so3_log(rotation_matrices)


def rot_transpose(mat):
    return mat.transpose(-1, -2)


def rot_mult(mat_1, mat_2):
    return torch.einsum("...ij,...jk->...ik", mat_1, mat_2)


def calc_rot_vf(mat_t, mat_1):
    return ops.so3_log(rot_mult(rot_transpose(mat_t), mat_1).contiguous())


def geodesic_t(t, mat, base_mat, rot_vf=None):
    """R_t = base * Exp(t * Log(base^T mat)).  t: scalar or [B,1,1]-like; mat, base_mat [B,L,3,3]."""
    if base_mat.shape != mat.shape:
        raise ValueError(f"Incompatible shapes: base_mat={base_mat.shape}, mat_t={mat.shape}")
    if rot_vf is not None:
        t = torch.as_tensor(t, device=mat.device, dtype=torch.float32)
        return torch.einsum("...ij,...jk->...ik", base_mat, ops.so3_exp((t * rot_vf).contiguous()))
    t = torch.as_tensor(t, device=mat.device, dtype=torch.float32)
    return ops.so3_geodesic(t, mat, base_mat)
