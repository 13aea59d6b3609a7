"""Error budget of the edge transition's split-precision GEMMs - CPU experiment, test infrastructure.

The tcgen05 kernel (pepflowww_b200/csrc/pf_edge_umma.cu) evaluates every pair-MLP GEMM as three fp16 products with fp32
accumulation: A_hi W_hi + A_lo W_hi + A_hi W_lo.  This script emulates that arithmetic inside the CPU oracle
(oracle/pepflow_oracle.py::edge_transition, reference models_con/ipa_pytorch.py:233-248) and measures what dropping one
cross term - per GEMM - does to the denoiser outputs after all six blocks at the headline residue count (256 + 15).

GEMMs of one edge transition as the kernel factors them (the per-residue parts P_i + Q_j / U_i + V_j are hoisted and
stay exact):   g1 = z W1z^T (K 64)   g2 = h1 W2^T (K 192)   g3z = z Wfz^T (K 64)   g3 = h2 Wf^T (K 192)
Modes per GEMM:  3 = hi*hi + lo*hi + hi*lo (shipped)   "a" = drop A_lo W_hi   "w" = drop A_hi W_lo   1 = hi*hi only

Writes profiles/r2_edge_error_budget.txt.   Run: python tests/experiments/edge_error_budget.py [L_pocket] [L_pep]
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pepflow_oracle as orc  # noqa: E402

torch.set_num_threads(os.cpu_count())


def split16(x):
    hi = x.half().float()
    lo = (x - hi).half().float()
    return hi, lo


def gemm(a, w, mode):
    """a [..., K] @ w[N, K]^T in emulated split precision."""
    if mode == "fp32":
        return a @ w.t()
    ah, al = split16(a)
    wh, wl = split16(w)
    y = ah @ wh.t()
    if mode in (3, "w"):
        y = y + al @ wh.t()
    if mode in (3, "a"):
        y = y + ah @ wl.t()
    return y


def make_edge_transition(modes):
    def edge_transition(sd, p, node, edge):
        B, L, _ = node.shape
        e = orc.linear(node, sd[p + "initial_embed.weight"], sd[p + "initial_embed.bias"])
        w1, b1 = sd[p + "trunk.0.weight"], sd[p + "trunk.0.bias"]
        w2, b2 = sd[p + "trunk.2.weight"], sd[p + "trunk.2.bias"]
        wf, bf = sd[p + "final_layer.weight"], sd[p + "final_layer.bias"]
        P = e @ w1[:, 64:128].t() + b1
        Q = e @ w1[:, 128:192].t()
        U = e @ wf[:, 64:128].t() + bf
        V = e @ wf[:, 128:192].t()
        h1 = torch.relu(gemm(edge, w1[:, :64], modes["g1"]) + P[:, :, None, :] + Q[:, None, :, :])
        h2 = torch.relu(gemm(h1, w2, modes["g2"]) + b2)
        y = gemm(h2, wf, modes["g3"]) + gemm(edge, wf[:, :64], modes["g3z"]) + U[:, :, None, :] + V[:, None, :, :]
        return orc.layer_norm(y, sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"])
    return edge_transition


def main():
    lr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    lp = int(sys.argv[2]) if len(sys.argv) > 2 else 15
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from pepflowww_b200.config import load_config
    from pepflowww_b200.flow_model import FlowModel
    from pepflowww_b200.pep_dataloader import synthetic_batch
    from pepflowww_b200.utils import deterministic_state_dict

    cfg, _ = load_config()
    sd = deterministic_state_dict(FlowModel(cfg.model).state_dict(), 114514)
    batch = synthetic_batch(1, lr, lp, seed=3)
    enc = orc.encode(sd, batch)
    B, L = batch["aa"].shape
    g = torch.Generator().manual_seed(1)
    gm = batch["generate_mask"]
    rot = torch.where(gm[..., None, None], orc.quat_to_rot(torch.nn.functional.normalize(torch.randn(B, L, 4, generator=g), dim=-1)), enc["rotmats_1"])
    tr = torch.where(gm[..., None], torch.randn(B, L, 3, generator=g), enc["trans_1"])
    ang = torch.where(gm[..., None], torch.rand(B, L, 5, generator=g) * 2 * math.pi, enc["angles_1"])
    seq = torch.where(gm, torch.randint(0, 20, (B, L), generator=g), enc["seqs_1"])
    t = torch.full((B, 1), 0.37)
    args = (sd, t, rot, tr, ang, seq, enc["node_embed"], enc["edge_embed"], gm.long(), batch["res_mask"].long())

    def run(modes):
        saved = orc.edge_transition
        orc.edge_transition = make_edge_transition(modes)
        try:
            return orc.ga_encoder_forward(*args)
        finally:
            orc.edge_transition = saved

    ref = orc.ga_encoder_forward(*args)

    def errs(out):
        e_rot = float((out[0] - ref[0]).abs().max())
        e_tr = float((out[1] - ref[1]).abs().max() / ref[1].abs().max())
        e_tr_abs = float((out[1] - ref[1]).abs().max())
        d = (out[2] - ref[2]).abs() % (2 * math.pi)
        e_ang = float(torch.minimum(d, 2 * math.pi - d).max())
        e_log = float((out[3] - ref[3]).abs().max())
        flips = int((out[3].argmax(-1) != ref[3].argmax(-1)).sum())
        return e_rot, e_tr, e_tr_abs, e_ang, e_log, flips

    names = ["g1", "g3z", "g2", "g3"]
    full = {k: 3 for k in names}
    cases = [("fp32 (hoisted form)", {k: "fp32" for k in names}), ("3-pass everywhere (shipped)", dict(full))]
    for k in names:
        for m in ("a", "w"):
            c = dict(full); c[k] = m
            cases.append((f"{k}: drop {'A_lo*W_hi' if m == 'a' else 'A_hi*W_lo'}", c))
    for m in ("a", "w"):
        cases.append((f"all four GEMMs: drop {'A_lo*W_hi' if m == 'a' else 'A_hi*W_lo'}", {k: m for k in names}))
    cases.append(("g1+g3z (z GEMMs, K 64): drop A_hi*W_lo", {"g1": "w", "g3z": "w", "g2": 3, "g3": 3}))
    cases.append(("g2+g3 (K 192): drop A_hi*W_lo", {"g1": 3, "g3z": 3, "g2": "w", "g3": "w"}))
    cases.append(("all four GEMMs: hi*hi only", {k: 1 for k in names}))

    lines = [f"# edge-transition split-precision error budget, oracle emulation, B=1, L={L} ({lr}+{lp}), 6 blocks (5 edge transitions)",
             "# errors against the plain fp32 oracle: rot = max |dR|, trans = max |dx| / max |x| (and absolute, Angstrom),",
             "# angles = max wrapped |da| (rad), logits = max |dl|, flips = residues whose argmax changes",
             f"# {'case':52s} {'rot':>9s} {'trans':>9s} {'trans_A':>9s} {'angles':>9s} {'logits':>9s} flips"]
    for name, modes in cases:
        e = errs(run(modes))
        lines.append(f"{name:54s} {e[0]:9.2e} {e[1]:9.2e} {e[2]:9.2e} {e[3]:9.2e} {e[4]:9.2e} {e[5]:3d}")
        print(lines[-1], flush=True)
    out = os.path.join(ROOT, "profiles", "r2_edge_error_budget.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
