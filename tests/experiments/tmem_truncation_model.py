"""CPU model of the tensor-memory accumulator rounding measured on the B200 (profiles/r2_gemm_error.txt: the accumulator
is truncated after every tcgen05.mma) - what it costs a K = 128 layer under both issue orders, and the K = 1536 IPA output
projection (12 chunks accumulated into one accumulator).  Assumes the 16 products of one MMA are summed exactly and the
sum is added to the fp32 accumulator with round-toward-zero.  Test infrastructure / analysis only.
    python tests/experiments/tmem_truncation_model.py     -> profiles/r2_tmem_truncation_model.txt"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rz32(x64):
    y = x64.astype(np.float32)
    over = np.abs(y.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(y, np.float32(0)), y).astype(np.float32)


def split16(x):
    hi = x.astype(np.float16).astype(np.float32)
    lo = (x - hi).astype(np.float16).astype(np.float32)
    return hi, lo


def gemm_model(x, w, order, chunk=128):
    """x [M,K], w [N,K]; K in chunks of 128 (8 MMAs of K=16 per product per chunk); returns fp32 [M,N]."""
    xh, xl = split16(x)
    wh, wl = split16(w)
    M, K = x.shape
    acc = np.zeros((M, w.shape[0]), np.float32)
    add = lambda acc, a, b: rz32(acc.astype(np.float64) + a.astype(np.float64) @ b.astype(np.float64).T)
    for c0 in range(0, K, chunk):
        ks = range(c0, min(c0 + chunk, K), 16)
        if order == 1:
            for k in ks:
                acc = add(acc, xl[:, k:k + 16], wh[:, k:k + 16])
                acc = add(acc, xh[:, k:k + 16], wl[:, k:k + 16])
            for k in ks:
                acc = add(acc, xh[:, k:k + 16], wh[:, k:k + 16])
        else:
            for k in ks:
                acc = add(acc, xl[:, k:k + 16], wh[:, k:k + 16])
                acc = add(acc, xh[:, k:k + 16], wl[:, k:k + 16])
                acc = add(acc, xh[:, k:k + 16], wh[:, k:k + 16])
    return acc


def main():
    rng = np.random.default_rng(0)
    lines = ["# tensor-memory accumulator model (round toward zero after every MMA), 3xFP16 split, against fp64",
             f"# {'case':44s} {'rms rel':>10s} {'mean signed':>12s}"]
    for name, K, positive in (("K=128 random sign", 128, False), ("K=128 positive", 128, True),
                              ("K=1536 random sign (IPA linear_out)", 1536, False), ("K=1536 positive", 1536, True)):
        x = rng.standard_normal((1024, K)).astype(np.float32)
        w = ((rng.random((128, K)) * 2 - 1) * 0.15).astype(np.float32)
        if positive:
            x, w = np.abs(x), np.abs(w)
        ref = x.astype(np.float64) @ w.astype(np.float64).T
        scale = np.abs(ref).mean()
        for order in (0, 1):
            y = gemm_model(x, w, order).astype(np.float64)
            lines.append(f"  {name + ', mma_order ' + str(order):44s} {np.sqrt(((y - ref) ** 2).mean()) / scale:10.2e} {(y - ref).mean() / scale:+12.2e}")
            print(lines[-1], flush=True)
        y = (x @ w.T).astype(np.float64)
        lines.append(f"  {name + ', fp32 sgemm (numpy)':44s} {np.sqrt(((y - ref) ** 2).mean()) / scale:10.2e} {(y - ref).mean() / scale:+12.2e}")
        print(lines[-1], flush=True)
    with open(os.path.join(ROOT, "profiles", "r2_tmem_truncation_model.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
