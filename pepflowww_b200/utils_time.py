"""Sinusoidal time embedding (reference: models_con/utils.py:60-72, called with max_positions=2056 from
models_con/ga.py:79-85).  The frequency table is computed on the host in fp32 with the same expression as
the reference so the CUDA feature-mix kernel multiplies by bit-identical frequencies."""
import math

import torch


def time_frequencies(embedding_dim=128, max_positions=2056):
    half = embedding_dim // 2
    step = math.log(max_positions) / (half - 1)
    return torch.exp(torch.arange(half, dtype=torch.float32) * -step)


_FREQ_CACHE = {}


def _frequencies_on(device, embedding_dim, max_positions):
    """Host-computed table, uploaded once per device (no host->device copy per call: CUDA-graph capturable)."""
    key = (str(device), embedding_dim, max_positions)
    if key not in _FREQ_CACHE:
        _FREQ_CACHE[key] = time_frequencies(embedding_dim, max_positions).to(device)
    return _FREQ_CACHE[key]


def get_time_embedding(timesteps, embedding_dim, max_positions=2000):
    assert len(timesteps.shape) == 1
    freqs = _frequencies_on(timesteps.device, embedding_dim, max_positions)
    arg = (timesteps * max_positions).float()[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1), mode="constant")
    return emb
