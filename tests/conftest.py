import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 114514


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def golden_state_dict_spec():
    spec = {}
    with open(os.path.join(GOLDEN, "state_dict_keys.txt")) as fh:
        for line in fh:
            parts = line.split()
            spec[parts[0]] = tuple(int(p) for p in parts[1:])
    return spec


@pytest.fixture(scope="session")
def state_dict():
    """Deterministic weights with the reference's keys/shapes (same filler make_golden.py used)."""
    from pepflowww_b200.utils import deterministic_state_dict
    spec = golden_state_dict_spec()
    proto = {k: torch.zeros(s) for k, s in spec.items()}
    nf = {"node_embedder.dihed_embed.freq_bands": 3, "edge_embedder.dihedral_embed.freq_bands": 3,
          "ga_encoder.angles_embedder.freq_bands": 12}
    for k, n in nf.items():
        proto[k] = torch.tensor([float(i + 1) for i in range(n)] + [1.0 / (i + 1) for i in range(n)])
    return deterministic_state_dict(proto, WEIGHT_SEED)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def circ_err(a, b):
    d = (a.double() - b.double()).abs() % (2 * np.pi)
    return float(torch.minimum(d, 2 * np.pi - d).max())
