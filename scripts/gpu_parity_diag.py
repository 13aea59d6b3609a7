"""Where does the denoiser's error against the CPU oracle come from at the headline residue count?  Runs GAEncoder.forward
at B=2, L=271 under different kernel-variant switches and prints the errors of each against the oracle; then the effect of
the repo's own encode (embedder kernels) on the first sampler iteration.  Test infrastructure (imports the oracle)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pepflow_oracle as orc  # noqa: E402
from pepflowww_b200 import _lib  # noqa: E402
from pepflowww_b200.config import load_config  # noqa: E402
from pepflowww_b200.flow_model import FlowModel  # noqa: E402
from pepflowww_b200.pep_dataloader import synthetic_batch  # noqa: E402
from pepflowww_b200.utils import deterministic_state_dict  # noqa: E402

GA_KEYS = ("t", "rotmats_t", "trans_t", "angles_t", "seqs_t", "node_embed", "edge_embed", "generate_mask", "res_mask")


def errs(out, ref):
    r = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    d = (out[2].double() - ref[2].double()).abs() % (2 * math.pi)
    return "rot %.2e trans %.2e ang %.2e logits %.2e" % (r(out[0], ref[0]), r(out[1], ref[1]),
                                                         float(torch.minimum(d, 2 * math.pi - d).max()), r(out[3], ref[3]))


def main():
    lr, lp = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 15)
    dev = torch.device("cuda:0")
    print("host CPU capability:", torch.backends.cpu.get_cpu_capability(), "threads", torch.get_num_threads())
    cfg, _ = load_config()
    model = FlowModel(cfg.model).eval()
    sd = deterministic_state_dict(model.state_dict(), 114514)
    model.load_state_dict(sd)
    model = model.to(dev)
    batch = synthetic_batch(2, lr, lp, seed=21)
    enc = orc.encode(sd, batch)
    B, L = batch["aa"].shape
    rng = np.random.default_rng(8)
    q = torch.from_numpy(rng.standard_normal((B, L, 4))).float()
    gm = batch["generate_mask"]
    inp = dict(t=torch.tensor([[0.21], [0.68]]),
               rotmats_t=torch.where(gm[..., None, None], orc.quat_to_rot(q / q.norm(dim=-1, keepdim=True)), enc["rotmats_1"]),
               trans_t=enc["trans_1"] + gm[..., None] * torch.from_numpy(rng.standard_normal((B, L, 3))).float(),
               angles_t=torch.from_numpy(rng.uniform(0, 2 * math.pi, (B, L, 5))).float(),
               seqs_t=torch.from_numpy(rng.integers(0, 20, (B, L))), node_embed=enc["node_embed"],
               edge_embed=enc["edge_embed"], generate_mask=gm.long(), res_mask=batch["res_mask"].long())
    ref = orc.ga_encoder_forward(sd, *[inp[k] for k in GA_KEYS])
    # double-precision oracle as the common yardstick
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    inp64 = {k: (v.double() if v.is_floating_point() else v) for k, v in inp.items()}
    try:
        ref64 = orc.ga_encoder_forward(sd64, *[inp64[k] for k in GA_KEYS])
        print("fp32 oracle vs fp64 oracle:           ", errs(ref, ref64))
    except Exception as e:      # noqa: BLE001
        ref64 = None
        print("fp64 oracle unavailable:", e)
    variants = [("fp32 everywhere", dict(edge_impl=0, gemm_impl=0, ipa_impl=0, chain_impl=0)),
                ("default", dict(edge_impl=2, gemm_impl=2, ipa_impl=4, chain_impl=1)),
                ("default, ipa fp32 (impl 0)", dict(edge_impl=2, gemm_impl=2, ipa_impl=0, chain_impl=1)),
                ("default, edge fp32 (impl 0)", dict(edge_impl=0, gemm_impl=2, ipa_impl=4, chain_impl=1)),
                ("default, unfused fp32 GEMMs", dict(edge_impl=2, gemm_impl=0, ipa_impl=4, chain_impl=0)),
                ("fp32 + ipa v4", dict(edge_impl=0, gemm_impl=0, ipa_impl=4, chain_impl=0)),
                ("fp32 + edge tcgen05", dict(edge_impl=2, gemm_impl=0, ipa_impl=0, chain_impl=0)),
                ("fp32 + chains", dict(edge_impl=0, gemm_impl=2, ipa_impl=0, chain_impl=1))]
    for name, opts in variants:
        for k, v in opts.items():
            _lib.set_option(k, v)
        with torch.no_grad():
            out = [o.cpu() for o in model.ga_encoder(*[inp[k].to(dev) for k in GA_KEYS])]
        line = f"{name:36s} vs fp32 oracle: {errs(out, ref)}"
        if ref64 is not None:
            line += f"   | vs fp64 oracle: {errs(out, [r.float() for r in ref64])}"
        print(line, flush=True)
    for k, v in dict(edge_impl=2, gemm_impl=2, ipa_impl=4, chain_impl=1).items():
        _lib.set_option(k, v)
    for name, opts in (("default, mma_order 0", dict()), ("fp32 + chains, mma_order 0", dict(edge_impl=0, ipa_impl=0))):
        for k, v in opts.items():
            _lib.set_option(k, v)
        _lib.set_option("mma_order", 0)
        with torch.no_grad():
            out = [o.cpu() for o in model.ga_encoder(*[inp[k].to(dev) for k in GA_KEYS])]
        print(f"{name:36s} vs fp32 oracle: {errs(out, ref)}", flush=True)
    _lib.set_option("mma_order", 1)
    for k, v in dict(edge_impl=2, gemm_impl=2, ipa_impl=4, chain_impl=1).items():
        _lib.set_option(k, v)
    for terms in (1, 3, 15):
        _lib.set_option("edge_terms", terms)
        with torch.no_grad():
            out = [o.cpu() for o in model.ga_encoder(*[inp[k].to(dev) for k in GA_KEYS])]
        print(f"default, edge_terms={terms:2d}                vs fp32 oracle: {errs(out, ref)}", flush=True)
    _lib.set_option("edge_terms", 0)
    # encode -> denoiser: the same denoiser input with the repo's own embedder outputs
    dbatch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    with torch.no_grad():
        own = model.encode(dbatch)
        inp_own = dict(inp, node_embed=own[4].cpu(), edge_embed=own[5].cpu())
        out = [o.cpu() for o in model.ga_encoder(*[inp_own[k].to(dev) for k in GA_KEYS])]
    r = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
    print("own encode vs oracle encode: node %.2e edge %.2e" % (r(own[4].cpu(), enc["node_embed"]), r(own[5].cpu(), enc["edge_embed"])))
    print(f"{'default, own encode':36s} vs fp32 oracle: {errs(out, ref)}")
    # the oracle itself with the kernel's embedder outputs: how much of that is the embedders' input perturbation
    ref_own = orc.ga_encoder_forward(sd, *[inp_own[k] for k in GA_KEYS])
    print(f"{'oracle on own-encode inputs':36s} vs fp32 oracle: {errs(ref_own, ref)}")
    print(f"{'default, own encode':36s} vs oracle on the same inputs: {errs(out, ref_own)}")


if __name__ == "__main__":
    main()
