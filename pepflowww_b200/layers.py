"""Host mirrors of pepflow/modules/common/layers.py: sample_from (:17-22), clampped_one_hot (:10-14),
AngularEncoding (:92-113)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def clampped_one_hot(x, num_classes):
    mask = (x >= 0) & (x < num_classes)
    x = x.clamp(min=0, max=num_classes - 1)
    return F.one_hot(x, num_classes) * mask[..., None]


def categorical_from_uniform(c, u):
    """Inverse-CDF draw from (c + 1e-8) given uniforms u in [0,1): the deterministic core the CUDA Euler
    kernels use (index = number of left-to-right prefix sums <= u * total)."""
    c = c + 1e-8
    cdf = torch.cumsum(c, dim=-1)
    thr = u[..., None] * cdf[..., -1:]
    return (cdf[..., :-1] <= thr).sum(dim=-1)


def sample_from(c, generator=None):
    """Categorical sample per residue from probabilities c [N, L, K] (reference: multinomial(c + 1e-8, 1))."""
    u = torch.rand(c.shape[:-1], device=c.device, dtype=c.dtype, generator=generator)
    return categorical_from_uniform(c, u)


class AngularEncoding(nn.Module):
    def __init__(self, num_funcs=3):
        super().__init__()
        self.num_funcs = num_funcs
        self.register_buffer("freq_bands", torch.FloatTensor(
            [i + 1 for i in range(num_funcs)] + [1.0 / (i + 1) for i in range(num_funcs)]))

    def get_out_dim(self, in_dim):
        return in_dim * (1 + 2 * 2 * self.num_funcs)

    def forward(self, x):
        shape = list(x.shape[:-1]) + [-1]
        x = x.unsqueeze(-1)
        code = torch.cat([x, torch.sin(x * self.freq_bands), torch.cos(x * self.freq_bands)], dim=-1)
        return code.reshape(shape)
