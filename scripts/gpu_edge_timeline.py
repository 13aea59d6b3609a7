"""Diagnostic: clock64() timeline of the tcgen05 edge-transition kernel (first 4 tiles of cluster 0).
python scripts/gpu_edge_timeline.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepflowww_b200 import _lib  # noqa: E402
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402

ROW_EV = ["start", "A0 staged", "acc1 ready", "h1 c0", "h1 c1", "h1 c2", "acc2 ready", "h2 c0", "h2 c1", "h2 c2",
          "acc3 ready", "tile end"]
MMA_EV = ["A0 ready", "L1 issued", "h1[0]", "h1[1]", "h1[2]", "h1[3]", "h1[4]", "h1[5]", "L2 issued", "h2[0]", "h2[1]",
          "h2[2]", "h2[3]", "h2[4]", "h2[5]", "L3 issued"]


def main():
    dev = torch.device("cuda:0")
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    model.load_state_dict(deterministic_state_dict(model.state_dict(), 114514))
    model = model.to(dev)
    et = model.ga_encoder.trunk["edge_transition_1"]
    B, L = 64, 271
    s = torch.randn(B, L, 128, device=dev)
    z = torch.randn(B, L, L, 64, device=dev)
    _lib.set_option("edge_impl", 2)
    with torch.no_grad():
        et(s, z)
        torch.cuda.synchronize()
        buf = torch.zeros(3 * 4 * 16, dtype=torch.int64, device=dev)
        lib = _lib.load()
        _lib.check(lib.pf_debug_buffer(ctypes.c_void_p(buf.data_ptr()), buf.numel() * 8))
        et(s, z)
        torch.cuda.synchronize()
        _lib.check(lib.pf_debug_buffer(None, 0))
    t = buf.cpu().view(3, 4, 16)
    t0 = int(t[0, 0, 0])
    for it in range(4):
        print(f"--- tile {it} (cycles since kernel's first stamp)")
        rows = []
        for actor, names in ((0, ROW_EV), (1, ROW_EV), (2, MMA_EV)):
            for e, n in enumerate(names):
                rows.append((int(t[actor, it, e]) - t0, ["rowgrp0", "rowgrp1", "mma    "][actor], n))
        for c, who, n in sorted(rows):
            print(f"  {c:8d}  {who}  {n}")


if __name__ == "__main__":
    main()
