// Node-level Linear layers: y = act(x W^T + b) (+ residual) (* rowmask).
// Replaces nn.Linear / ipa_pytorch.Linear (models_con/ipa_pytorch.py:116-181) on the hot path.
//
// Two sm_100a variants behind pf_linear (option "gemm_impl"):
//   0: fp32 CUDA-core SGEMM (64x64x16 tiles, 4x4 register micro-tiles, register-prefetch double buffer)
//   1: 3xFP16 split-precision tensor-core GEMM (mma.sync m16n8k16, fp32 accumulate): each fp32
//      operand is split into fp16 hi + fp16 lo and the product is hi*hi + lo*hi + hi*lo, which keeps
//      ~22 significand bits - single-pass TF32/BF16 fails the 1e-4 parity bar (SURVEY.md finding 5).
#include "pf_common.cuh"
#include "pf_split.cuh"

namespace pf {

// ------------------------------------------------------------------------------------------------
// fp32 SGEMM:  A [M,K] row-major, W [N,K] row-major (both K-contiguous), Y [M,N].
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

template <bool VEC>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       const float* __restrict__ bias,
                                                       const float* __restrict__ residual,
                                                       const float* __restrict__ rowmask, float* __restrict__ Y,
                                                       int M, int K, int N, int ldw, int act) {
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Ws[2][BK][BN + PADM];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int lrow = tid >> 2;          // 0..63: tile row this thread loads
  const int lk = (tid & 3) * 4;       // k offset (4 consecutive k) this thread loads
  const int ty = tid >> 4, tx = tid & 15;  // 16x16 thread grid, 4x4 outputs each

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rw[4];
  auto gload = [&](int k0) {
    const int am = m0 + lrow, wn = n0 + lrow, k = k0 + lk;
    if (VEC) {
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
      if (am < M && k < K) va = *reinterpret_cast<const float4*>(A + (size_t)am * K + k);
      if (wn < N && k < K) vw = *reinterpret_cast<const float4*>(W + (size_t)wn * ldw + k);
      ra[0] = va.x; ra[1] = va.y; ra[2] = va.z; ra[3] = va.w;
      rw[0] = vw.x; rw[1] = vw.y; rw[2] = vw.z; rw[3] = vw.w;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ra[e] = (am < M && k + e < K) ? A[(size_t)am * K + k + e] : 0.f;
        rw[e] = (wn < N && k + e < K) ? W[(size_t)wn * ldw + k + e] : 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[buf][lk + e][lrow] = ra[e];
      Ws[buf][lk + e][lrow] = rw[e];
    }
  };

  const int nk = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const float rm = rowmask ? rowmask[m] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == 1) v = fmaxf(v, 0.f);
      if (residual) v += residual[(size_t)m * N + n];
      Y[(size_t)m * N + n] = v * rm;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 3xFP16 tensor-core GEMM (mma.sync.m16n8k16).  CTA tile 128(M) x 64(N), BK = 32, 8 warps, each warp
// 16 rows x 64 cols (8 n-tiles).  Operands are split to fp16 hi/lo while they are staged into smem.
// ------------------------------------------------------------------------------------------------
constexpr int TM = 128, TN = 64, TK = 32, SKP = TK + 8;  // smem row = 40 halves = 80 B (16 B-aligned, ldmatrix conflict-free)

template <bool VEC>
__global__ void __launch_bounds__(256) f16x3_gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ residual,
                                                             const float* __restrict__ rowmask,
                                                             float* __restrict__ Y, int M, int K, int N, int ldw, int act) {
  __shared__ __align__(16) __half sAh[TM][SKP], sAl[TM][SKP];
  __shared__ __align__(16) __half sWh[TN][SKP], sWl[TN][SKP];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + TK - 1) / TK;
  for (int kt = 0; kt < nk; ++kt) {
    const int k0 = kt * TK;
    // stage A: 128 rows x 32 k = 1024 float4 -> 4 per thread; W: 64 x 32 = 512 float4 -> 2 per thread
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = tid + it * 256;
      const int r = idx >> 3, kq = (idx & 7) * 4;
      const int m = m0 + r, k = k0 + kq;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < M) {
        if (VEC) {
          if (k < K) {
            const float4 f = *reinterpret_cast<const float4*>(A + (size_t)m * K + k);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (k + e < K) v[e] = A[(size_t)m * K + k + e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) split_one(v[e], sAh[r][kq + e], sAl[r][kq + e]);
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256;
      const int r = idx >> 3, kq = (idx & 7) * 4;
      const int n = n0 + r, k = k0 + kq;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (n < N) {
        if (VEC) {
          if (k < K) {
            const float4 f = *reinterpret_cast<const float4*>(W + (size_t)n * ldw + k);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (k + e < K) v[e] = W[(size_t)n * ldw + k + e];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) split_one(v[e], sWh[r][kq + e], sWl[r][kq + e]);
    }
    __syncthreads();
    // Two-level accumulation: the tensor core adds products into its fp32 accumulator with truncation, which
    // biases long chains (measured: 5x the fp32 error over K = 629..1536).  Each K-tile therefore accumulates
    // into a zeroed chunk (6 chained MMAs) that is added to the master accumulator with a rounded FADD.
    float chunk[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) chunk[i][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < TK; ks += 16) {
      uint32_t ah[4], al[4];
      const int r0 = warp * 16 + g;
      ah[0] = *reinterpret_cast<const uint32_t*>(&sAh[r0][ks + 2 * t]);
      ah[1] = *reinterpret_cast<const uint32_t*>(&sAh[r0 + 8][ks + 2 * t]);
      ah[2] = *reinterpret_cast<const uint32_t*>(&sAh[r0][ks + 2 * t + 8]);
      ah[3] = *reinterpret_cast<const uint32_t*>(&sAh[r0 + 8][ks + 2 * t + 8]);
      al[0] = *reinterpret_cast<const uint32_t*>(&sAl[r0][ks + 2 * t]);
      al[1] = *reinterpret_cast<const uint32_t*>(&sAl[r0 + 8][ks + 2 * t]);
      al[2] = *reinterpret_cast<const uint32_t*>(&sAl[r0][ks + 2 * t + 8]);
      al[3] = *reinterpret_cast<const uint32_t*>(&sAl[r0 + 8][ks + 2 * t + 8]);
#pragma unroll
      for (int n0 = 0; n0 < 8; n0 += 4) {              // 4 independent accumulator chains per group
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int nr = (n0 + q) * 8 + g;
          bh[q][0] = *reinterpret_cast<const uint32_t*>(&sWh[nr][ks + 2 * t]);
          bh[q][1] = *reinterpret_cast<const uint32_t*>(&sWh[nr][ks + 2 * t + 8]);
          bl[q][0] = *reinterpret_cast<const uint32_t*>(&sWl[nr][ks + 2 * t]);
          bl[q][1] = *reinterpret_cast<const uint32_t*>(&sWl[nr][ks + 2 * t + 8]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) mma16816(chunk[n0 + q], al, bh[q][0], bh[q][1]);   // small terms first
#pragma unroll
        for (int q = 0; q < 4; ++q) mma16816(chunk[n0 + q], ah, bl[q][0], bl[q][1]);
#pragma unroll
        for (int q = 0; q < 4; ++q) mma16816(chunk[n0 + q], ah, bh[q][0], bh[q][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] += chunk[i][j];
    __syncthreads();
  }

  // epilogue: c0,c1 -> (row g, cols 2t,2t+1); c2,c3 -> (row g+8, ...)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int m = m0 + warp * 16 + g + half * 8;
    if (m >= M) continue;
    const float rm = rowmask ? rowmask[m] : 1.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n0 + nt * 8 + 2 * t + e;
        if (n >= N) continue;
        float v = acc[nt][half * 2 + e] + (bias ? bias[n] : 0.f);
        if (act == 1) v = fmaxf(v, 0.f);
        if (residual) v += residual[(size_t)m * N + n];
        Y[(size_t)m * N + n] = v * rm;
      }
    }
  }
}

int launch_linear_full(const float* x, const float* w, int ldw, const float* bias, const float* residual,
                       const float* rowmask, float* y, int M, int K, int N, int act, cudaStream_t st) {
  if (M == 0 || N == 0) return PF_OK;
  const bool vec = (K % 4 == 0) && (ldw % 4 == 0) && aligned16(x) && aligned16(w);
  if (opt_gemm_impl() >= 1) {
    dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
    if (vec) f16x3_gemm_tn_kernel<true><<<grid, 256, 0, st>>>(x, w, bias, residual, rowmask, y, M, K, N, ldw, act);
    else f16x3_gemm_tn_kernel<false><<<grid, 256, 0, st>>>(x, w, bias, residual, rowmask, y, M, K, N, ldw, act);
  } else {
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    if (vec) sgemm_tn_kernel<true><<<grid, 256, 0, st>>>(x, w, bias, residual, rowmask, y, M, K, N, ldw, act);
    else sgemm_tn_kernel<false><<<grid, 256, 0, st>>>(x, w, bias, residual, rowmask, y, M, K, N, ldw, act);
  }
  PF_CHECK_LAUNCH();
  return PF_OK;
}

// With a caller-provided workspace the K = 128 layers run on the tcgen05 GEMM ("gemm_impl" = 2); everything else
// (and every call without workspace) takes the mma.sync / fp32 kernels above.
size_t linear_workspace_bytes(int N) { return gemm_umma_pack_bytes(N) + 256; }

bool linear_umma_eligible(int K, int N, bool has_residual) { return K == 128 && N % 4 == 0 && !has_residual; }

int launch_linear_ws(const float* x, const float* w, int ldw, const float* bias, const float* residual,
                     const float* rowmask, float* y, int M, int K, int N, int act, void* ws, size_t ws_bytes,
                     cudaStream_t st, const void* prepacked) {
  if (M == 0 || N == 0) return PF_OK;
  const bool shape_ok = opt_gemm_impl() == 2 && linear_umma_eligible(K, N, residual != nullptr) && aligned16(x) &&
                        aligned16(y) && ldw % 2 == 0;
  if (shape_ok && prepacked && aligned16(prepacked))
    return launch_linear_umma(x, w, ldw, bias, rowmask, y, M, N, act, prepacked, true, st);
  if (shape_ok && ws && ws_bytes >= gemm_umma_pack_bytes(N) && aligned16(ws))
    return launch_linear_umma(x, w, ldw, bias, rowmask, y, M, N, act, ws, false, st);
  return launch_linear_full(x, w, ldw, bias, residual, rowmask, y, M, K, N, act, st);
}

int launch_linear(const float* x, const float* w, const float* bias, const float* residual, const float* rowmask,
                  float* y, int M, int K, int N, int act, cudaStream_t st) {
  return launch_linear_full(x, w, K, bias, residual, rowmask, y, M, K, N, act, st);
}

// W is a column block of a wider matrix: row stride ldw (used for the per-residue parts of the edge MLP).
int launch_linear_ld(const float* x, const float* w, int ldw, const float* bias, float* y, int M, int K, int N,
                     cudaStream_t st) {
  return launch_linear_full(x, w, ldw, bias, nullptr, nullptr, y, M, K, N, 0, st);
}

}  // namespace pf

extern "C" size_t pf_linear_workspace_bytes(int N) { return pf::linear_workspace_bytes(N); }

extern "C" int pf_linear_ws(const float* x, const float* w, const float* bias, const float* residual,
                            const float* rowmask, float* y, int M, int K, int N, int act, void* workspace,
                            size_t workspace_bytes, void* stream) {
  PF_REQUIRE(x && w && y, PF_ERR_NULL_POINTER);
  PF_REQUIRE(M >= 0 && K > 0 && N > 0 && (act == 0 || act == 1), PF_ERR_BAD_SHAPE);
  return pf::launch_linear_ws(x, w, K, bias, residual, rowmask, y, M, K, N, act, workspace, workspace_bytes,
                              pf::as_stream(stream));
}

extern "C" int pf_linear(const float* x, const float* w, const float* bias, const float* residual,
                         const float* rowmask, float* y, int M, int K, int N, int act, void* stream) {
  PF_REQUIRE(x && w && y, PF_ERR_NULL_POINTER);
  PF_REQUIRE(M >= 0 && K > 0 && N > 0 && (act == 0 || act == 1), PF_ERR_BAD_SHAPE);
  return pf::launch_linear(x, w, bias, residual, rowmask, y, M, K, N, act, pf::as_stream(stream));
}
