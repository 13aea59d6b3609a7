"""Minimal Rigid / Rotation containers with the reference's method names
(openfold/utils/rigid_utils.py:289-850, 853-1468), backed by the CUDA kernels.  Only the subset the hot path
touches: create_rigid (data/utils.py:42-44), get_rots/get_trans/get_rot_mats/get_quats, apply,
invert_apply, compose_q_update_vec."""
import torch

from . import ops


class Rotation:
    def __init__(self, rot_mats=None, quats=None, normalize_quats=True):
        if (rot_mats is None) == (quats is None):
            raise ValueError("Exactly one input argument must be specified")
        if (rot_mats is not None and rot_mats.shape[-2:] != (3, 3)) or (quats is not None and quats.shape[-1] != 4):
            raise ValueError("Incorrectly shaped rotation matrix or quaternion")
        if quats is not None:
            quats = quats.type(torch.float32)
            if normalize_quats:
                quats = quats / torch.linalg.norm(quats, dim=-1, keepdim=True)
        if rot_mats is not None:
            rot_mats = rot_mats.type(torch.float32)
        self._rot_mats, self._quats = rot_mats, quats

    def get_rot_mats(self):
        return self._rot_mats if self._rot_mats is not None else ops.quat_to_rot(self._quats)

    def get_quats(self):
        if self._quats is not None:
            return self._quats
        # rot -> quat shares the rigid-update kernel (zero update leaves the frame unchanged)
        z3 = torch.zeros(*self._rot_mats.shape[:-2], 3, device=self._rot_mats.device)
        z6 = torch.zeros(*self._rot_mats.shape[:-2], 6, device=self._rot_mats.device)
        m = torch.ones(self._rot_mats.shape[:-2], device=self._rot_mats.device)
        q, _, _ = ops.rigid_update(None, self._rot_mats, z3, z6, m)
        return q

    def apply(self, pts):
        return torch.einsum("...ij,...j->...i", self.get_rot_mats(), pts)

    def invert_apply(self, pts):
        return torch.einsum("...ji,...j->...i", self.get_rot_mats(), pts)


class Rigid:
    def __init__(self, rots, trans):
        if rots is None or trans is None:
            raise ValueError("rots and trans are required")
        self._rots, self._trans = rots, trans.type(torch.float32)

    def get_rots(self):
        return self._rots

    def get_trans(self):
        return self._trans

    def apply(self, pts):
        return self._rots.apply(pts) + self._trans

    def invert_apply(self, pts):
        return self._rots.invert_apply(pts - self._trans)

    def compose_q_update_vec(self, q_update_vec, update_mask=None):
        """q' = normalize(q + m q (x) (0, v)); t' = t + m R(q) u   - one kernel (pf_rigid_update)."""
        shape = self._trans.shape[:-1]
        m = torch.ones(shape, device=self._trans.device) if update_mask is None else update_mask.reshape(shape).float()
        q, r, t = ops.rigid_update(self._rots._quats, self._rots._rot_mats if self._rots._quats is None else None,
                                   self._trans, q_update_vec, m)
        new = Rigid(Rotation(quats=q, normalize_quats=False), t)
        new._rots._cached_rot = r
        return new


def create_rigid(rots, trans):
    return Rigid(Rotation(rot_mats=rots), trans)
