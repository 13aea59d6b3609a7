"""Why the IPA point term was reformulated (finding 3) - CPU experiment, test infrastructure.

Evaluates InvariantPointAttention (oracle, models_con/ipa_pytorch.py:316-484) at the benchmark geometry with the
point-distance logits  -1/2 c_h sum_p |q_p - k_p|^2  computed three ways in fp32:
  direct    differences first (the reference)
  expanded  c_h q.k - 1/2 c_h |k|^2 (- row constant): the round-1 tensor-core form, every product / partial sum rounded to fp32
  split     -P/2 c_h |t_i - t_j|^2 exactly + c_h (t_i.B_j + A_i.t_j + sum_p a.b) - c_h t_j.B_j - 1/2 c_h sum|b|^2: the round-2 form
and reports the error of the logits and of the module output against an fp64 evaluation, with the complex at its own
position (peptide centroid at the origin, |t| up to ~45 A) and translated by 100 A.

Writes profiles/r2_ipa_point_term_forms.txt.      python tests/experiments/ipa_point_term_forms.py
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pepflow_oracle as orc  # noqa: E402
import benchdata  # noqa: E402

H, C, PQ, PV = 8, 128, 8, 12


def ipa(sd, p, s, z, R, t, mask, form, dt):
    """orc.ipa_forward with a selectable point-term form; everything except that term in dtype `dt`."""
    cast = lambda x: x.to(dt)
    B, L, _ = s.shape
    lin = lambda name, x: cast(x) @ cast(sd[p + name + ".weight"]).t() + cast(sd[p + name + ".bias"])
    q = lin("linear_q", s).view(B, L, H, C)
    kv = lin("linear_kv", s).view(B, L, H, 2 * C)
    k, v = kv[..., :C], kv[..., C:]
    Rd, td = cast(R), cast(t)

    def local(name, n):
        return lin(name, s).view(B, L, 3, H, n).permute(0, 1, 3, 4, 2)

    q_loc, kv_loc = local("linear_q_points", PQ), local("linear_kv_points", PQ + PV)
    rot = lambda x: torch.einsum("blij,blhnj->blhni", Rd, x)
    a_pts, b_all = rot(q_loc), rot(kv_loc)                      # rotated only
    b_pts, v_pts = b_all[..., :PQ, :], b_all[..., PQ:, :] + td[:, :, None, None, :]
    hw = torch.nn.functional.softplus(cast(sd[p + "head_weights"])) * math.sqrt(1.0 / (3 * (PQ * 9.0 / 2)))   # [H]
    q_pts, k_pts = a_pts + td[:, :, None, None, :], b_pts + td[:, :, None, None, :]
    if form == "direct":
        d2 = ((q_pts[:, :, None] - k_pts[:, None, :]) ** 2).sum((-1, -2))                       # [B,L,L,H]
        pt = -0.5 * d2 * hw
    elif form == "expanded":
        qk = torch.einsum("bihpx,bjhpx->bijh", q_pts, k_pts)                                      # products of size |t|^2
        kk = (k_pts ** 2).sum((-1, -2))                                                             # [B,L,H]
        pt = hw * qk - 0.5 * hw * kk[:, None]
        pt = pt - pt.amax(2, keepdim=True).detach() * 0                                           # row constant irrelevant
    else:
        A, Bs = a_pts.sum(3), b_pts.sum(3)                                                        # [B,L,H,3]
        D = ((td[:, :, None] - td[:, None, :]) ** 2).sum(-1)                                      # exact per pair
        cross = torch.einsum("bix,bjhx->bijh", td, Bs) + torch.einsum("bihx,bjx->bijh", A, td) + \
            torch.einsum("bihpx,bjhpx->bijh", a_pts, b_pts)
        kb = -hw * (torch.einsum("bjx,bjhx->bjh", td, Bs) + 0.5 * (b_pts ** 2).sum((-1, -2)))
        pt = -0.5 * PQ * hw * D[..., None] + hw * cross + kb[:, None]
    bias = cast(z) @ cast(sd[p + "linear_b.weight"]).t() + cast(sd[p + "linear_b.bias"])
    logits = torch.einsum("bihc,bjhc->bijh", q, k) * math.sqrt(1.0 / (3 * C)) + math.sqrt(1.0 / 3) * bias + pt
    m = cast(mask)
    logits = logits + (1e5 * (m[:, :, None] * m[:, None, :] - 1))[..., None]
    logits = logits - logits.amax(2, keepdim=True)                                                # softmax-invariant shift
    a = torch.softmax(logits, dim=2)                                                              # over j
    o = torch.einsum("bijh,bjhc->bihc", a, v).reshape(B, L, H * C)
    o_pt = torch.einsum("bijh,bjhpx->bihpx", a, v_pts) - td[:, :, None, None, :]
    o_pt = torch.einsum("blji,blhpj->blhpi", Rd, o_pt)
    o_norm = torch.sqrt((o_pt ** 2).sum(-1) + 1e-8).reshape(B, L, H * PV)
    o_pt = o_pt.reshape(B, L, H * PV, 3)
    pair_z = cast(z) @ cast(sd[p + "down_z.weight"]).t() + cast(sd[p + "down_z.bias"])
    o_pair = torch.einsum("bijh,bijc->bihc", a, pair_z).reshape(B, L, -1)
    feats = torch.cat([o, o_pt[..., 0], o_pt[..., 1], o_pt[..., 2], o_norm, o_pair], dim=-1)
    return feats @ cast(sd[p + "linear_out.weight"]).t() + cast(sd[p + "linear_out.bias"]), logits


def main():
    torch.set_num_threads(os.cpu_count())
    sd = benchdata.reference_state_dict(114514)
    batch = benchdata.synthetic_batch(1, 256, 15, seed=21)
    enc = orc.encode(sd, batch)
    B, L = batch["aa"].shape
    g = torch.Generator().manual_seed(3)
    s = torch.randn(B, L, 128, generator=g)
    mask = torch.ones(B, L)
    p = "ga_encoder.trunk.ipa_1."
    lines = ["# IPA point-term forms in fp32 against fp64 (oracle restatement, 1 complex of 256 + 15 residues, ipa_1, deterministic weights)",
             "# logits: max |d logit| over pairs whose attention weight exceeds 1e-6; output: max |d out| / max |out|",
             f"# {'shift':>8s} {'form':>9s} {'max |t|':>8s} {'logits':>10s} {'output':>10s}"]
    for shift in (0.0, 100.0):
        t = enc["trans_1"] + shift
        ref, ref_logits = ipa(sd, p, s, enc["edge_embed"], enc["rotmats_1"], t, mask, "direct", torch.float64)
        w = torch.softmax(ref_logits, dim=2) > 1e-6
        for form in ("direct", "expanded", "split"):
            out, lg = ipa(sd, p, s, enc["edge_embed"], enc["rotmats_1"], t, mask, form, torch.float32)
            e_l = float(((lg.double() - ref_logits).abs() * w).max())
            e_o = float((out.double() - ref).abs().max() / ref.abs().max())
            lines.append(f"  {shift:8.0f} {form:>9s} {float(t.abs().max()):8.1f} {e_l:10.2e} {e_o:10.2e}")
            print(lines[-1], flush=True)
    with open(os.path.join(ROOT, "profiles", "r2_ipa_point_term_forms.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
