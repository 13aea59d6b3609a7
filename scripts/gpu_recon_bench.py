"""Times the post-sampling reconstruction kernels (csrc/pf_recon.cu) with CUDA events: the last trajectory entry of the
bench batch (64 x 271 residues) and a whole 200-step trajectory (3.47 M residues), against their algorithmic bytes.
    python scripts/gpu_recon_bench.py"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pepflowww_b200 import constants, ops  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    T = constants.rigid_tables(dev)
    peak = 6552.6
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    out = {}
    ncu = "--ncu" in sys.argv      # under ncu: one launch of each kernel at trajectory scale, no warm-up loop
    global timed
    if ncu:
        timed = lambda fn, reps=1: (fn(), torch.cuda.synchronize(), 1.0)[2]
    for tag, B in ((("trajectory_200x64x271", 64 * 200),) if ncu else
                   (("final_step_64x271", 64), ("trajectory_200x64x271", 64 * 200))):
        L = 271
        q = torch.randn(B, L, 4, device=dev)
        q = q / q.norm(dim=-1, keepdim=True)
        a, b, c, d = q.unbind(-1)
        R = torch.stack([a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c), 2 * (b * c + a * d),
                         a * a - b * b + c * c - d * d, 2 * (c * d - a * b), 2 * (b * d - a * c), 2 * (c * d + a * b),
                         a * a - b * b - c * c + d * d], -1).reshape(B, L, 3, 3).contiguous()
        t = torch.randn(B, L, 3, device=dev) * 10
        ang = torch.rand(B, L, 5, device=dev) * 2 * math.pi
        aa = torch.randint(0, 20, (B, L), device=dev)
        res_nb = torch.arange(1, L + 1, device=dev).repeat(B, 1)
        chain_nb = torch.zeros(B, L, dtype=torch.long, device=dev)
        mask = torch.ones(B, L, dtype=torch.bool, device=dev)
        n = B * L
        ms_pos = timed(lambda: ops.full_atom_reconstruction(R, t, ang, aa, T, want_frames=False, want_mask=True))
        ms_all = timed(lambda: ops.full_atom_reconstruction(R, t, ang, aa, T, want_frames=True, want_mask=False))
        ms_bb = timed(lambda: ops.reconstruct_backbone(R, t, aa, chain_nb, res_nb, mask, T))
        pos14 = ops.full_atom_reconstruction(R, t, ang, aa, T, want_frames=False)[0]
        ms_tor = timed(lambda: ops.torsion_angles(pos14, aa, T))
        by_tor = n * (42 * 4 + 8 + 5 * 4 + 5)
        by_pos = n * ((9 + 3 + 5) * 4 + 8 + 42 * 4 + 15)          # frames, torsions, types in; pos14 + mask out
        by_all = n * ((9 + 3 + 5) * 4 + 8 + (42 + 54 + 18) * 4)    # + the six frames out
        by_bb = n * ((9 + 3) * 4 + 3 * 8 + 1 + 12 * 4)
        out[tag] = {"residues": n,
                    "pos14+mask": {"ms": ms_pos, "GB/s": by_pos / ms_pos / 1e6, "frac_hbm": by_pos / ms_pos / 1e6 / peak},
                    "pos14+frames": {"ms": ms_all, "GB/s": by_all / ms_all / 1e6, "frac_hbm": by_all / ms_all / 1e6 / peak},
                    "torsion_angles": {"ms": ms_tor, "GB/s": by_tor / ms_tor / 1e6, "frac_hbm": by_tor / ms_tor / 1e6 / peak},
                    "backbone": {"ms": ms_bb, "GB/s": by_bb / ms_bb / 1e6, "frac_hbm": by_bb / ms_bb / 1e6 / peak}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
