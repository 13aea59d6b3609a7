"""Diagnostic (not a test): per-stage error of the CUDA kernels against the oracle, chaining the granular
C-ABI ops exactly like pf_ga_encoder_forward does.  python tests/gpu_diag.py [edge_impl gemm_impl]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pepflow_oracle as orc  # noqa: E402
from pepflowww_b200 import _lib, ops  # noqa: E402
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402
from pepflowww_b200.utils_time import time_frequencies  # noqa: E402
from tests.conftest import load_golden  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu().double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def main():
    edge_impl, gemm_impl = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 0)
    ipa_impl = int(os.environ.get("IPA_IMPL", "4"))
    _lib.set_option("edge_impl", edge_impl)
    _lib.set_option("gemm_impl", gemm_impl)
    _lib.set_option("ipa_impl", ipa_impl)
    print("ipa_impl", ipa_impl)
    dev = torch.device("cuda:0")
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    sd = deterministic_state_dict(model.state_dict(), 114514)
    model.load_state_dict(sd)
    model = model.to(dev)
    tag = sys.argv[3] if len(sys.argv) > 3 else "ga_encoder_a"
    if tag.startswith("synth"):   # synth:<pocket>:<peptide>  - embedder outputs of a synthetic batch, like smoke()
        from pepflowww_b200.pep_dataloader import synthetic_batch
        _, lr, lp = tag.split(":")
        batch = synthetic_batch(2, int(lr), int(lp), seed=5)
        enc = orc.encode(sd, batch)
        B, L = batch["aa"].shape
        gen = torch.Generator().manual_seed(0)
        gm = batch["generate_mask"]
        R0 = orc.quat_to_rot(torch.nn.functional.normalize(torch.randn(B, L, 4, generator=gen), dim=-1))
        g = {"t": torch.full((B, 1), 0.01), "rotmats_t": torch.where(gm[..., None, None], R0, enc["rotmats_1"]),
             "trans_t": torch.where(gm[..., None], torch.randn(B, L, 3, generator=gen), enc["trans_1"]),
             "angles_t": torch.where(gm[..., None], torch.rand(B, L, 5, generator=gen) * 2 * math.pi, enc["angles_1"]),
             "seqs_t": torch.where(gm, torch.randint(0, 20, (B, L), generator=gen), enc["seqs_1"]),
             "node_embed": enc["node_embed"], "edge_embed": enc["edge_embed"], "generate_mask": gm.long(),
             "res_mask": batch["res_mask"].long()}
    else:
        g = load_golden(tag)
    keys = ("t", "rotmats_t", "trans_t", "angles_t", "seqs_t", "node_embed", "edge_embed", "generate_mask", "res_mask")
    trace = []
    ref = orc.ga_encoder_forward(sd, *[g[k] for k in keys], trace=trace)
    tr = dict(trace)
    ga = model.ga_encoder
    D = lambda x: x.to(dev)
    m = D(g["res_mask"]).float()
    B, L = g["seqs_t"].shape
    print(f"edge_impl={edge_impl} gemm_impl={gemm_impl}  B={B} L={L}")
    with torch.no_grad():
        x = ops.mix_features(D(g["node_embed"]), ga.current_seq_embedder.weight, D(g["seqs_t"]), D(g["t"]),
                             time_frequencies().to(dev), D(g["angles_t"]), ga.angles_embedder.freq_bands)
        h = ops.linear(x, ga.res_feat_mixer[0].weight, ga.res_feat_mixer[0].bias, act=1)
        s = ops.linear(h, ga.res_feat_mixer[2].weight, ga.res_feat_mixer[2].bias, rowmask=m.reshape(-1))
        print("mix        %.2e" % rel(s, tr["mix"]))
        rot, trans, quat = D(g["rotmats_t"]), D(g["trans_t"]), None
        z = D(g["edge_embed"])
        for b in range(6):
            T = ga.trunk
            ipa = T[f"ipa_{b}"]
            w, bias = ipa.packed_projection()
            proj = ops.linear(s, w, bias)
            pts = ops.ipa_points(proj, rot, trans)
            feats = ops.ipa_attention(proj, pts, z, ipa.linear_b.weight, ipa.linear_b.bias, ipa.down_z.weight,
                                      ipa.down_z.bias, ipa.scaled_head_weights(), rot, trans, m)
            ipa_out = ops.linear(feats, ipa.linear_out.weight, ipa.linear_out.bias, rowmask=m.reshape(-1))
            e_ipa = rel(ipa_out, tr[f"ipa_{b}"])
            # teacher-forced: IPA from the oracle's inputs of this block
            s = ops.add_layernorm(s, ipa_out, T[f"ipa_ln_{b}"].weight, T[f"ipa_ln_{b}"].bias)
            y = s
            for layer in T[f"seq_tfmr_{b}"].layers:
                qkv = ops.linear(y, layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias)
                ctx = ops.seq_attention(qkv, m)
                a = ops.linear(ctx, layer.self_attn.out_proj.weight, layer.self_attn.out_proj.bias)
                y1 = ops.add_layernorm(y, a, layer.norm1.weight, layer.norm1.bias)
                f = ops.linear(ops.linear(y1, layer.linear1.weight, layer.linear1.bias, act=1), layer.linear2.weight,
                               layer.linear2.bias)
                y = ops.add_layernorm(y1, f, layer.norm2.weight, layer.norm2.bias)
            valid = g["res_mask"].bool()
            e_tf = rel(y.cpu()[valid], tr[f"tfmr_{b}"][valid])
            s = ops.linear(y, T[f"post_tfmr_{b}"].weight, T[f"post_tfmr_{b}"].bias, residual=s)
            nt = T[f"node_transition_{b}"]
            hh = ops.linear(ops.linear(s, nt.linear_1.weight, nt.linear_1.bias, act=1), nt.linear_2.weight,
                            nt.linear_2.bias, act=1)
            hh = ops.linear(hh, nt.linear_3.weight, nt.linear_3.bias)
            s = ops.add_layernorm(s, hh, nt.ln.weight, nt.ln.bias, rowmask=m.reshape(-1))
            e_node = rel(s, tr[f"node_{b}"])
            upd = ops.linear(s, T[f"bb_update_{b}"].linear.weight, T[f"bb_update_{b}"].linear.bias)
            quat, rot, trans = ops.rigid_update(quat, rot if quat is None else None, trans, upd, m)
            e_rot, e_tr = rel(rot, tr[f"rot_{b}"]), rel(trans, tr[f"trans_{b}"])
            e_z = float("nan")
            if b < 5:
                z = T[f"edge_transition_{b}"](s, z, edge_mask_rows=m)
                e_z = rel(z, tr[f"z_{b}"])
            print(f"block {b}: ipa {e_ipa:.2e} tfmr {e_tf:.2e} node {e_node:.2e} rot {e_rot:.2e} trans {e_tr:.2e} z {e_z:.2e}")
        out = ga(*[D(g[k]) for k in keys])
        print("composite vs oracle: rot %.2e trans %.2e logits %.2e" % (rel(out[0], ref[0]), rel(out[1], ref[1]), rel(out[3], ref[3])))
        print("composite vs chain : rot %.2e trans %.2e" % (rel(out[0], rot.cpu()), rel(out[1], trans.cpu())))
        # teacher-forced seams on oracle inputs
        for b in (0, 3):
            s_in = D(tr["mix"] if b == 0 else tr[f"node_{b-1}"])
            z_in = D(g["edge_embed"] if b == 0 else tr[f"z_{b-1}"])
            r_in = D(g["rotmats_t"] if b == 0 else tr[f"rot_{b-1}"])
            t_in = D(g["trans_t"] if b == 0 else tr[f"trans_{b-1}"])
            ipa = ga.trunk[f"ipa_{b}"]
            w, bias = ipa.packed_projection()
            proj = ops.linear(s_in, w, bias)
            pts = ops.ipa_points(proj, r_in, t_in)
            feats = ops.ipa_attention(proj, pts, z_in, ipa.linear_b.weight, ipa.linear_b.bias, ipa.down_z.weight,
                                      ipa.down_z.bias, ipa.scaled_head_weights(), r_in, t_in, m)
            ipa_out = ops.linear(feats, ipa.linear_out.weight, ipa.linear_out.bias, rowmask=m.reshape(-1))
            print(f"teacher-forced ipa_{b}: {rel(ipa_out, tr[f'ipa_{b}']):.2e}  (|ipa|max {float(tr[f'ipa_{b}'].abs().max()):.2f})")


if __name__ == "__main__":
    main()
