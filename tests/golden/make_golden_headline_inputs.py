"""Seeds and the denoiser-input builder shared by make_golden_headline.py (reference side) and
tests/test_oracle_golden.py (oracle side) - no dependency on the reference."""
import math

import numpy as np
import torch

WEIGHT_SEED, DATA_SEED, NOISE_SEED = 114514, 21, 8


def headline_inputs(enc, batch):
    """The denoiser input both sides build from encoder outputs `enc` (dict) - also imported by the tests."""
    B, L = batch["aa"].shape
    rng = np.random.default_rng(NOISE_SEED)
    q = torch.from_numpy(rng.standard_normal((B, L, 4))).float()
    q = q / q.norm(dim=-1, keepdim=True)
    a, b, c, d = q.unbind(-1)
    rot = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
                       2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
                       2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], -1).reshape(B, L, 3, 3)
    gm = batch["generate_mask"]
    return dict(t=torch.tensor([[0.21], [0.68]]),
                rotmats_t=torch.where(gm[..., None, None], rot, enc["rotmats_1"]),
                trans_t=enc["trans_1"] + gm[..., None] * torch.from_numpy(rng.standard_normal((B, L, 3))).float(),
                angles_t=torch.from_numpy(rng.uniform(0, 2 * math.pi, (B, L, 5))).float(),
                seqs_t=torch.from_numpy(rng.integers(0, 20, (B, L))), node_embed=enc["node_embed"],
                edge_embed=enc["edge_embed"], generate_mask=gm.long(), res_mask=batch["res_mask"].long())
