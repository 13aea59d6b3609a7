"""Torus T^d maps with the reference's names (models_con/torus.py:5-26).  tor_geodesic_t runs the CUDA
kernel (pf_tor_geodesic); the one-liners used only by the training loss stay torch elementwise ops."""
import math

import torch

from . import ops


def tor_expmap(x, u):
    return (x + u) % (2 * math.pi)


def tor_logmap(x, y):
    return torch.atan2(torch.sin(y - x), torch.cos(y - x))


def tor_projx(x):
    return x % (2 * math.pi)


def tor_random_uniform(*size, dtype=None, device=None, generator=None):
    return torch.rand(*size, dtype=dtype, device=device, generator=generator) * 2 * math.pi


def tor_geodesic_t(t, angles_1, angles_0):
    """(angles_0 + t * log_{angles_0}(angles_1)) mod 2pi; t broadcast over leading dims ([B,1,1] or scalar)."""
    if not angles_1.is_cuda:
        raise RuntimeError("tor_geodesic_t: CUDA tensors required (no CPU fallback)")
    t = torch.as_tensor(t, device=angles_1.device, dtype=torch.float32)
    return ops.tor_geodesic(t, angles_1, angles_0)
