"""Diagnostic (not a test): edge-transition variants against each other + timing.
python scripts/gpu_edge_check.py [B L]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepflowww_b200 import _lib  # noqa: E402
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def main():
    dev = torch.device("cuda:0")
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), 114514))
    model = model.to(dev)
    et = model.ga_encoder.trunk["edge_transition_1"]
    shapes = [(2, 30), (1, 16), (3, 37), (2, 140), (2, 271)]
    if len(sys.argv) > 2:
        shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
    for B, L in shapes:
        g = torch.Generator().manual_seed(2)
        s = torch.randn(B, L, 128, generator=g).to(dev)
        z = torch.randn(B, L, L, 64, generator=g).to(dev)
        outs = {}
        for impl in (0, 1, 2):
            _lib.set_option("edge_impl", impl)
            with torch.no_grad():
                outs[impl] = et(s, z)
            torch.cuda.synchronize()
        print(f"B={B} L={L}: v1 vs v0 {rel(outs[1], outs[0]):.2e}   umma vs v0 {rel(outs[2], outs[0]):.2e}", flush=True)
        bad = (outs[2] - outs[0]).abs().amax(dim=-1)
        if rel(outs[2], outs[0]) > 1e-4:
            idx = (bad > 1e-3).nonzero()
            print("  bad pairs:", idx.shape[0], "of", bad.numel(), "first:", idx[:8].tolist())
            b0, i0, j0 = idx[0].tolist()
            print("  got", outs[2][b0, i0, j0, :8].tolist())
            print("  ref", outs[0][b0, i0, j0, :8].tolist())
    B, L = 64, 271
    g = torch.Generator().manual_seed(3)
    s = torch.randn(B, L, 128, generator=g).to(dev)
    z = torch.randn(B, L, L, 64, device=dev)
    for impl in (1, 2):
        _lib.set_option("edge_impl", impl)
        with torch.no_grad():
            for _ in range(2):
                et(s, z)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                et(s, z)
            e1.record()
            torch.cuda.synchronize()
        print(f"impl {impl}: {e0.elapsed_time(e1) / 5:.3f} ms per edge transition (B=64, L=271, incl. small GEMMs)", flush=True)


if __name__ == "__main__":
    main()
