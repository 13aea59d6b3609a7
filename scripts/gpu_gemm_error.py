"""Error statistics of the three GEMM paths (fp32 FMA, 3xFP16 mma.sync, 3xFP16 tcgen05) against fp64: rms and mean signed
relative error, random-sign and all-positive operands (a biased mean on positive operands = the accumulator is truncated,
not rounded), both MMA issue orders of the tcgen05 kernel.  Test infrastructure."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pepflowww_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
M, K, N = 8192, 128, 128
for kind in ("random sign", "positive"):
    x = torch.randn(M, K, generator=g)
    w = (torch.rand(N, K, generator=g) * 2 - 1) * 0.15
    if kind == "positive":
        x, w = x.abs(), w.abs()
    ref = x.double() @ w.double().t()
    scale = ref.abs().mean()
    cpu32 = (x @ w.t()).double()
    print(f"[{kind}] torch CPU fp32 sgemm          rms {float(((cpu32 - ref) ** 2).mean().sqrt() / scale):.2e} mean {float((cpu32 - ref).mean() / scale):+.2e}")
    for impl, order in ((0, 0), (1, 0), (2, 0), (2, 1)):
        _lib.set_option("gemm_impl", impl)
        _lib.set_option("mma_order", order)
        y = ops.linear(x.to(dev), w.to(dev)).cpu().double()
        print(f"[{kind}] gemm_impl {impl} mma_order {order}        rms {float(((y - ref) ** 2).mean().sqrt() / scale):.2e} mean {float((y - ref).mean() / scale):+.2e}")
_lib.set_option("gemm_impl", 2)
_lib.set_option("mma_order", 1)
