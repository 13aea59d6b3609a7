// K8: edge transition (models_con/ipa_pytorch.py:233-248, mask from models_con/ga.py:118).
//
//   e = W_init s + b_init (128 -> 64 per residue);  x_ij = [z_ij | e_i | e_j]  (192)
//   h1 = relu(W1 x + b1); h2 = relu(W2 h1 + b2); y = W_f (h2 + x) + b_f; z'_ij = LN_64(y) * m_i m_j
//
// 93.7 % of the denoiser's FLOPs (SURVEY.md finding 4).  Two sm_100a variants ("edge_impl"):
//   0: fp32 CUDA-core kernel that follows the formula literally (x materialised in shared memory).
//   1: 3xFP16 split-precision tensor-core kernel (mma.sync m16n8k16, fp32 accumulate).  The per-residue
//      parts of W1 x and W_f x are hoisted out of the pair loop (P_i + Q_j, U_i + V_j; verified 2.7e-6 in
//      SURVEY.md App. F), the MLP chain stays in registers (accumulator fragments are re-used as the next
//      GEMM's A fragments), W2 is resident in shared memory and W1z/W_f stream through a cp.async ring.
//      z is read once and written once (in place allowed).
#include "pf_common.cuh"
#include "pf_split.cuh"

namespace pf {

// =================================================================================================
// variant 0: fp32, literal
// =================================================================================================
constexpr int E0_ROWS = 64;      // pair rows per CTA
constexpr int E0_XS = 196;       // smem row stride of the activation tiles (192 + 4)
constexpr int E0_WS = 196;       // smem row stride of a weight chunk [16][N]

struct EdgeArgs {
  const float* e;      // [B*L, 64]
  const float* z_in;   // [B, L*L, 64]
  const float* w1; const float* b1; const float* w2; const float* b2; const float* wf; const float* bf;
  const float* ln_g; const float* ln_b; const float* mask;
  float* z_out;
  int B, L;
};

// Y[64][N] (+)= X[64][192] * W[N][192]^T, thread (rg = warp, cg = lane): rows rg*8..+7, cols cg + 32 m.
template <int NM>
__device__ __forceinline__ void e0_gemm(const float* __restrict__ W, const float* X, float* Wc, float (&acc)[8][NM],
                                        int tid) {
  const int rg = tid >> 5, cg = tid & 31;
  constexpr int N = NM * 32;
  for (int kc = 0; kc < 192; kc += 16) {
    __syncthreads();
    for (int idx = tid; idx < N * 16; idx += 256) {  // W chunk -> Wc[k][n]
      const int n = idx >> 4, k = idx & 15;
      Wc[k * E0_WS + n] = W[(size_t)n * 192 + kc + k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float wv[NM];
#pragma unroll
      for (int m = 0; m < NM; ++m) wv[m] = Wc[k * E0_WS + cg + 32 * m];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float xv = X[(rg * 8 + r) * E0_XS + kc + k];
#pragma unroll
        for (int m = 0; m < NM; ++m) acc[r][m] = fmaf(xv, wv[m], acc[r][m]);
      }
    }
  }
}

__global__ void __launch_bounds__(256) edge_transition_v0_kernel(EdgeArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                       // [64][196]  x, later h2 + x
  float* Hb = X + E0_ROWS * E0_XS;       // [64][196]  h1
  float* Wc = Hb + E0_ROWS * E0_XS;      // [16][196]
  const int L = a.L, LL = L * L;
  const int b = blockIdx.y, r0 = blockIdx.x * E0_ROWS;
  const int tid = threadIdx.x, rg = tid >> 5, cg = tid & 31;
  const size_t rowb = (size_t)b * L;

  for (int idx = tid; idx < E0_ROWS * 192; idx += 256) {
    const int r = idx / 192, c = idx % 192;
    const int pr = r0 + r;
    float v = 0.f;
    if (pr < LL) {
      const int i = pr / L, j = pr % L;
      if (c < 64) v = a.z_in[((size_t)b * LL + pr) * CZ + c];
      else if (c < 128) v = a.e[(rowb + i) * 64 + (c - 64)];
      else v = a.e[(rowb + j) * 64 + (c - 128)];
    }
    X[r * E0_XS + c] = v;
  }
  // e0_gemm starts with a __syncthreads()
  float acc[8][6];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int m = 0; m < 6; ++m) acc[r][m] = a.b1[cg + 32 * m];
  e0_gemm<6>(a.w1, X, Wc, acc, tid);
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int m = 0; m < 6; ++m) Hb[(rg * 8 + r) * E0_XS + cg + 32 * m] = fmaxf(acc[r][m], 0.f);
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int m = 0; m < 6; ++m) acc[r][m] = a.b2[cg + 32 * m];
  e0_gemm<6>(a.w2, Hb, Wc, acc, tid);
  // X <- relu(h2) + x  (each thread touches only its own elements)
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int m = 0; m < 6; ++m) X[(rg * 8 + r) * E0_XS + cg + 32 * m] += fmaxf(acc[r][m], 0.f);
  float y[8][2];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int m = 0; m < 2; ++m) y[r][m] = a.bf[cg + 32 * m];
  e0_gemm<2>(a.wf, X, Wc, y, tid);
  // LayerNorm over the 64 outputs of each row (held by the 32 lanes of this warp), pair mask, store
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int pr = r0 + rg * 8 + r;
    const float mu = warp_sum(y[r][0] + y[r][1]) * (1.0f / 64.0f);
    const float d0 = y[r][0] - mu, d1 = y[r][1] - mu;
    const float rstd = 1.0f / sqrtf(warp_sum(d0 * d0 + d1 * d1) * (1.0f / 64.0f) + 1e-5f);
    if (pr < LL) {
      const int i = pr / L, j = pr % L;
      const float pm = a.mask[rowb + i] * a.mask[rowb + j];
      float* out = a.z_out + ((size_t)b * LL + pr) * CZ;
      out[cg] = (d0 * rstd * a.ln_g[cg] + a.ln_b[cg]) * pm;
      out[cg + 32] = (d1 * rstd * a.ln_g[cg + 32] + a.ln_b[cg + 32]) * pm;
    }
  }
}

// =================================================================================================
// variant 1: 3xFP16 tensor cores
// =================================================================================================
// Packed weight fragment layout: for k-step ks (16 k), n-tile nt (8 n), lane = g*4+t:
//   16 bytes = { hi(k=2t,2t+1), hi(k=2t+8,2t+9), lo(2t,2t+1), lo(2t+8,2t+9) } of row n = 8 nt + g
// i.e. exactly the B fragments of mma.m16n8k16 (hi, lo) for that lane -> one conflict-free LDS.128.
constexpr int E1_NT_WIDE = 24, E1_NT_OUT = 8;
constexpr int E1_KS_Z = 4, E1_KS_H = 12;
constexpr int E1_SLAB_WIDE = E1_NT_WIDE * 32 * 16;  // 12288 B: one k-step of a 192-wide layer
constexpr int E1_SLAB_OUT = E1_NT_OUT * 32 * 16;    //  4096 B: one k-step of the 64-wide layer
constexpr int E1_W2_BYTES = E1_KS_H * E1_SLAB_WIDE; // 147456 B resident
constexpr int E1_STAGE = E1_SLAB_WIDE;              // ring stage size
constexpr int E1_NSTAGE = 3;
// streamed stages per tile: 4 x W1z k-step | 4 x (3 k-steps of Wf) | 2 x (2 k-steps of Wfz)
constexpr int E1_STAGES_PER_TILE = 10;
constexpr int E1_STREAM_BYTES = 4 * E1_SLAB_WIDE + 12 * E1_SLAB_OUT + 4 * E1_SLAB_OUT;  // 114688
constexpr int E1_ROWS = 128;
constexpr int E1_G = 4;                              // n-tiles interleaved per MMA group

// 3xFP16 product of one A fragment with G consecutive n-tiles of packed B fragments.  The three MMAs of an
// n-tile form a dependent chain on its accumulator, so G independent chains are interleaved to cover the
// mma.sync latency (measured: strictly sequential chains ran at ~15 cycles per MMA per scheduler).
template <int G, int NACC>
__device__ __forceinline__ void mma3_group(float (&acc)[NACC][4], int nt0, const uint32_t (&ah)[4],
                                           const uint32_t (&al)[4], const uint4* wp, int lane) {
  uint4 w[G];
#pragma unroll
  for (int q = 0; q < G; ++q) w[q] = wp[q * 32 + lane];
#pragma unroll
  for (int q = 0; q < G; ++q) mma16816(acc[nt0 + q], al, w[q].x, w[q].y);
#pragma unroll
  for (int q = 0; q < G; ++q) mma16816(acc[nt0 + q], ah, w[q].z, w[q].w);
#pragma unroll
  for (int q = 0; q < G; ++q) mma16816(acc[nt0 + q], ah, w[q].x, w[q].y);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Repack kernel: fp32 weights -> packed fp16 hi/lo fragments.
//   w2pack  : W2  [192 n][192 k]            -> [12 ks][24 nt][32][16 B]
//   stream  : W1z = W1[:, 0:64]  (4 ks x 24 nt) | Wf [64 n][192 k] (12 ks x 8 nt) | Wfz = Wf[:, 0:64] (4 ks x 8 nt)
__global__ void edge_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                         const float* __restrict__ wf, uint4* __restrict__ w2pack,
                                         uint4* __restrict__ stream) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_w2 = E1_W2_BYTES / 16, n_w1z = 4 * E1_SLAB_WIDE / 16, n_wf = 12 * E1_SLAB_OUT / 16,
            n_wfz = 4 * E1_SLAB_OUT / 16;
  const float* W;
  int rel, nts;
  uint4* dst;
  if (idx < n_w2) { W = w2; rel = idx; nts = E1_NT_WIDE; dst = w2pack + idx; }
  else if (idx < n_w2 + n_w1z) { W = w1; rel = idx - n_w2; nts = E1_NT_WIDE; dst = stream + rel; }
  else if (idx < n_w2 + n_w1z + n_wf) { W = wf; rel = idx - n_w2 - n_w1z; nts = E1_NT_OUT; dst = stream + n_w1z + rel; }
  else if (idx < n_w2 + n_w1z + n_wf + n_wfz) { W = wf; rel = idx - n_w2 - n_w1z - n_wf; nts = E1_NT_OUT; dst = stream + n_w1z + n_wf + rel; }
  else return;
  const int lane = rel & 31, nt = (rel >> 5) % nts, ks = (rel >> 5) / nts;
  const int g = lane >> 2, t = lane & 3;
  const float* row = W + (size_t)(nt * 8 + g) * 192 + ks * 16 + 2 * t;
  uint4 o;
  split_pair(row[0], row[1], o.x, o.z);
  split_pair(row[8], row[9], o.y, o.w);
  *dst = o;
}

struct Edge1Args {
  const float* z_in;     // [B, L*L, 64]
  const float* P;        // [B*L, 192]  W1[:,64:128] e_i + b1
  const float* Q;        // [B*L, 192]  W1[:,128:192] e_j
  const float* U;        // [B*L, 64]   Wf[:,64:128] e_i + bf
  const float* V;        // [B*L, 64]   Wf[:,128:192] e_j
  const float* b2;       // [192]
  const float* ln_g; const float* ln_b; const float* mask;
  const uint4* w2pack; const uint4* stream;
  float* z_out;
  int B, L, tiles_per_complex, total_tiles;
};

__device__ __forceinline__ int e1_stage_bytes(int s) { return s < 8 ? E1_SLAB_WIDE : 2 * E1_SLAB_OUT; }
__device__ __forceinline__ int e1_stage_offset(int s) {
  return s < 8 ? s * E1_SLAB_WIDE : 8 * E1_SLAB_WIDE + (s - 8) * 2 * E1_SLAB_OUT;
}

__global__ void __launch_bounds__(256, 1) edge_transition_v1_kernel(Edge1Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* sW2 = reinterpret_cast<uint4*>(smem_raw);                                   // resident W2 fragments
  unsigned char* ring = smem_raw + E1_W2_BYTES;                                      // E1_NSTAGE x E1_STAGE
  float* sB2 = reinterpret_cast<float*>(ring + E1_NSTAGE * E1_STAGE);                // [192]
  float* sG = sB2 + 192;                                                             // [64] ln gamma
  float* sBt = sG + 64;                                                              // [64] ln beta
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int L = a.L, LL = L * L;

  // resident loads
  for (int i = tid; i < E1_W2_BYTES / 16; i += 256) cp_async16(sW2 + i, a.w2pack + i);
  for (int i = tid; i < 192; i += 256) sB2[i] = a.b2[i];
  if (tid < 64) { sG[tid] = a.ln_g[tid]; sBt[tid] = a.ln_b[tid]; }

  const int my_tiles = (a.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_stages = my_tiles * E1_STAGES_PER_TILE;
  int issued = 0, consumed = 0;
  auto issue_stage = [&]() {
    if (issued < total_stages) {
      const int s = issued % E1_STAGES_PER_TILE;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.stream) + e1_stage_offset(s);
      unsigned char* dst = ring + (issued % E1_NSTAGE) * E1_STAGE;
      const int n16 = e1_stage_bytes(s) / 16;
      for (int i = tid; i < n16; i += 256) cp_async16(dst + i * 16, src + i * 16);
    }
    cp_async_commit();
    ++issued;
  };
  // first group carries the resident W2 as well
  issue_stage();
  issue_stage();
  // acquire the next streamed stage: returns its smem base
  auto acquire = [&]() -> const uint4* {
    cp_async_wait<E1_NSTAGE - 2>();
    __syncthreads();           // stage `consumed` landed for all threads; stage consumed-1 is free
    issue_stage();             // refill the buffer of stage consumed-1
    const uint4* p = reinterpret_cast<const uint4*>(ring + (consumed % E1_NSTAGE) * E1_STAGE);
    ++consumed;
    return p;
  };

  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int b = tile / a.tiles_per_complex;
    const int r0 = (tile % a.tiles_per_complex) * E1_ROWS + warp * 16;
    const int pr_lo = r0 + g, pr_hi = r0 + g + 8;
    const bool ok_lo = pr_lo < LL, ok_hi = pr_hi < LL;
    const size_t rowb = (size_t)b * L;
    const int i_lo = ok_lo ? pr_lo / L : 0, j_lo = ok_lo ? pr_lo % L : 0;
    const int i_hi = ok_hi ? pr_hi / L : 0, j_hi = ok_hi ? pr_hi % L : 0;
    const float* z_lo = a.z_in + ((size_t)b * LL + (ok_lo ? pr_lo : 0)) * CZ;
    const float* z_hi = a.z_in + ((size_t)b * LL + (ok_hi ? pr_hi : 0)) * CZ;

    uint32_t ah[12][4], al[12][4];   // A fragments (hi / lo) of the current layer input
    float acc[24][4];

    // ---- A fragments of z (K = 64 -> k-steps 0..3)
#pragma unroll
    for (int ks = 0; ks < E1_KS_Z; ++ks) {
      const int c = ks * 16 + 2 * t;
      const float2 l0 = ok_lo ? *reinterpret_cast<const float2*>(z_lo + c) : make_float2(0.f, 0.f);
      const float2 h0 = ok_hi ? *reinterpret_cast<const float2*>(z_hi + c) : make_float2(0.f, 0.f);
      const float2 l1 = ok_lo ? *reinterpret_cast<const float2*>(z_lo + c + 8) : make_float2(0.f, 0.f);
      const float2 h1 = ok_hi ? *reinterpret_cast<const float2*>(z_hi + c + 8) : make_float2(0.f, 0.f);
      split_pair(l0.x, l0.y, ah[ks][0], al[ks][0]);
      split_pair(h0.x, h0.y, ah[ks][1], al[ks][1]);
      split_pair(l1.x, l1.y, ah[ks][2], al[ks][2]);
      split_pair(h1.x, h1.y, ah[ks][3], al[ks][3]);
    }
    // ---- layer 1 accumulators start at P_i + Q_j (b1 folded into P)
    {
      const float* P_lo = a.P + (rowb + i_lo) * 192; const float* Q_lo = a.Q + (rowb + j_lo) * 192;
      const float* P_hi = a.P + (rowb + i_hi) * 192; const float* Q_hi = a.Q + (rowb + j_hi) * 192;
#pragma unroll
      for (int nt = 0; nt < 24; ++nt) {
        const int n = nt * 8 + 2 * t;
        const float2 p0 = *reinterpret_cast<const float2*>(P_lo + n), q0 = *reinterpret_cast<const float2*>(Q_lo + n);
        const float2 p1 = *reinterpret_cast<const float2*>(P_hi + n), q1 = *reinterpret_cast<const float2*>(Q_hi + n);
        acc[nt][0] = p0.x + q0.x; acc[nt][1] = p0.y + q0.y;
        acc[nt][2] = p1.x + q1.x; acc[nt][3] = p1.y + q1.y;
      }
    }
    // ---- GEMM 1: z (K=64) x W1z, streamed one k-step per stage
#pragma unroll
    for (int ks = 0; ks < E1_KS_Z; ++ks) {
      const uint4* wp = acquire();
#pragma unroll
      for (int nt = 0; nt < 24; nt += E1_G) mma3_group<E1_G>(acc, nt, ah[ks], al[ks], wp + nt * 32, lane);
    }
    // ---- h1 = relu(.) -> A fragments (C fragment of n-tiles 2ks, 2ks+1 == A fragment of k-step ks)
#pragma unroll
    for (int ks = 0; ks < E1_KS_H; ++ks) {
      split_pair(fmaxf(acc[2 * ks][0], 0.f), fmaxf(acc[2 * ks][1], 0.f), ah[ks][0], al[ks][0]);
      split_pair(fmaxf(acc[2 * ks][2], 0.f), fmaxf(acc[2 * ks][3], 0.f), ah[ks][1], al[ks][1]);
      split_pair(fmaxf(acc[2 * ks + 1][0], 0.f), fmaxf(acc[2 * ks + 1][1], 0.f), ah[ks][2], al[ks][2]);
      split_pair(fmaxf(acc[2 * ks + 1][2], 0.f), fmaxf(acc[2 * ks + 1][3], 0.f), ah[ks][3], al[ks][3]);
    }
    // ---- GEMM 2: h1 (K=192) x W2 (resident)
#pragma unroll
    for (int nt = 0; nt < 24; ++nt) {
      const float2 bv = *reinterpret_cast<const float2*>(sB2 + nt * 8 + 2 * t);
      acc[nt][0] = bv.x; acc[nt][1] = bv.y; acc[nt][2] = bv.x; acc[nt][3] = bv.y;
    }
#pragma unroll
    for (int ks = 0; ks < E1_KS_H; ++ks) {
      const uint4* wp = sW2 + ks * (E1_NT_WIDE * 32);
#pragma unroll
      for (int nt = 0; nt < 24; nt += E1_G) mma3_group<E1_G>(acc, nt, ah[ks], al[ks], wp + nt * 32, lane);
    }
    // ---- h2 = relu(.) -> A fragments
#pragma unroll
    for (int ks = 0; ks < E1_KS_H; ++ks) {
      split_pair(fmaxf(acc[2 * ks][0], 0.f), fmaxf(acc[2 * ks][1], 0.f), ah[ks][0], al[ks][0]);
      split_pair(fmaxf(acc[2 * ks][2], 0.f), fmaxf(acc[2 * ks][3], 0.f), ah[ks][1], al[ks][1]);
      split_pair(fmaxf(acc[2 * ks + 1][0], 0.f), fmaxf(acc[2 * ks + 1][1], 0.f), ah[ks][2], al[ks][2]);
      split_pair(fmaxf(acc[2 * ks + 1][2], 0.f), fmaxf(acc[2 * ks + 1][3], 0.f), ah[ks][3], al[ks][3]);
    }
    // ---- GEMM 3: y = Wf h2 + Wfz z + U_i + V_j (bf folded into U)
    float y[8][4];
    {
      const float* U_lo = a.U + (rowb + i_lo) * 64; const float* V_lo = a.V + (rowb + j_lo) * 64;
      const float* U_hi = a.U + (rowb + i_hi) * 64; const float* V_hi = a.V + (rowb + j_hi) * 64;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int n = nt * 8 + 2 * t;
        const float2 u0 = *reinterpret_cast<const float2*>(U_lo + n), v0 = *reinterpret_cast<const float2*>(V_lo + n);
        const float2 u1 = *reinterpret_cast<const float2*>(U_hi + n), v1 = *reinterpret_cast<const float2*>(V_hi + n);
        y[nt][0] = u0.x + v0.x; y[nt][1] = u0.y + v0.y;
        y[nt][2] = u1.x + v1.x; y[nt][3] = u1.y + v1.y;
      }
    }
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const uint4* wp = acquire();
#pragma unroll
      for (int k3 = 0; k3 < 3; ++k3) {
        const int ks = st * 3 + k3;
#pragma unroll
        for (int nt = 0; nt < 8; nt += E1_G)
          mma3_group<E1_G>(y, nt, ah[ks], al[ks], wp + (k3 * E1_NT_OUT + nt) * 32, lane);
      }
    }
    // z again (registers were recycled): A fragments for the residual path
#pragma unroll
    for (int ks = 0; ks < E1_KS_Z; ++ks) {
      const int c = ks * 16 + 2 * t;
      const float2 l0 = ok_lo ? *reinterpret_cast<const float2*>(z_lo + c) : make_float2(0.f, 0.f);
      const float2 h0 = ok_hi ? *reinterpret_cast<const float2*>(z_hi + c) : make_float2(0.f, 0.f);
      const float2 l1 = ok_lo ? *reinterpret_cast<const float2*>(z_lo + c + 8) : make_float2(0.f, 0.f);
      const float2 h1 = ok_hi ? *reinterpret_cast<const float2*>(z_hi + c + 8) : make_float2(0.f, 0.f);
      split_pair(l0.x, l0.y, ah[ks][0], al[ks][0]);
      split_pair(h0.x, h0.y, ah[ks][1], al[ks][1]);
      split_pair(l1.x, l1.y, ah[ks][2], al[ks][2]);
      split_pair(h1.x, h1.y, ah[ks][3], al[ks][3]);
    }
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const uint4* wp = acquire();
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int ks = st * 2 + k2;
#pragma unroll
        for (int nt = 0; nt < 8; nt += E1_G)
          mma3_group<E1_G>(y, nt, ah[ks], al[ks], wp + (k2 * E1_NT_OUT + nt) * 32, lane);
      }
    }
    // ---- LayerNorm over 64 outputs: a row lives in the 4 lanes of a quad (16 values each)
    float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s_lo += y[nt][0] + y[nt][1]; s_hi += y[nt][2] + y[nt][3]; }
    s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
    s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
    const float mu_lo = s_lo * (1.0f / 64.0f), mu_hi = s_hi * (1.0f / 64.0f);
    float q_lo = 0.f, q_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float d0 = y[nt][0] - mu_lo, d1 = y[nt][1] - mu_lo, d2 = y[nt][2] - mu_hi, d3 = y[nt][3] - mu_hi;
      q_lo += d0 * d0 + d1 * d1; q_hi += d2 * d2 + d3 * d3;
    }
    q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 1); q_lo += __shfl_xor_sync(0xffffffffu, q_lo, 2);
    q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 1); q_hi += __shfl_xor_sync(0xffffffffu, q_hi, 2);
    const float rs_lo = 1.0f / sqrtf(q_lo * (1.0f / 64.0f) + 1e-5f), rs_hi = 1.0f / sqrtf(q_hi * (1.0f / 64.0f) + 1e-5f);
    const float pm_lo = ok_lo ? a.mask[rowb + i_lo] * a.mask[rowb + j_lo] : 0.f;
    const float pm_hi = ok_hi ? a.mask[rowb + i_hi] * a.mask[rowb + j_hi] : 0.f;
    float* o_lo = a.z_out + ((size_t)b * LL + (ok_lo ? pr_lo : 0)) * CZ;
    float* o_hi = a.z_out + ((size_t)b * LL + (ok_hi ? pr_hi : 0)) * CZ;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int n = nt * 8 + 2 * t;
      const float g0 = sG[n], g1 = sG[n + 1], c0 = sBt[n], c1 = sBt[n + 1];
      if (ok_lo)
        *reinterpret_cast<float2*>(o_lo + n) = make_float2(((y[nt][0] - mu_lo) * rs_lo * g0 + c0) * pm_lo,
                                                           ((y[nt][1] - mu_lo) * rs_lo * g1 + c1) * pm_lo);
      if (ok_hi)
        *reinterpret_cast<float2*>(o_hi + n) = make_float2(((y[nt][2] - mu_hi) * rs_hi * g0 + c0) * pm_hi,
                                                           ((y[nt][3] - mu_hi) * rs_hi * g1 + c1) * pm_hi);
    }
  }
  cp_async_wait<0>();
}

constexpr size_t E1_SMEM = E1_W2_BYTES + E1_NSTAGE * E1_STAGE + (192 + 64 + 64) * sizeof(float);
constexpr size_t E0_SMEM = (size_t)(2 * E0_ROWS * E0_XS + 16 * E0_WS) * sizeof(float);

void edge_kernels_init() {
  cudaFuncSetAttribute(edge_transition_v0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E0_SMEM);
  cudaFuncSetAttribute(edge_transition_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E1_SMEM);
  edge_umma_init();
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t edge_workspace_bytes(int B, int L) {
  const size_t M = (size_t)B * L;
  return align256(M * 64 * 4) + 2 * align256(M * 192 * 4) + 2 * align256(M * 64 * 4) + align256(E1_W2_BYTES) +
         align256(E1_STREAM_BYTES) + align256(edge_umma_pack_bytes(B, L));
}

// The per-residue terms of the hoisted first / last layer inside the edge-transition workspace
void edge_term_buffers(void* workspace, int B, int L, float** P, float** Q, float** U, float** V) {
  const size_t M = (size_t)B * L;
  unsigned char* ws = static_cast<unsigned char*>(workspace) + align256(M * 64 * 4);
  *P = reinterpret_cast<float*>(ws); ws += align256(M * 192 * 4);
  *Q = reinterpret_cast<float*>(ws); ws += align256(M * 192 * 4);
  *U = reinterpret_cast<float*>(ws); ws += align256(M * 64 * 4);
  *V = reinterpret_cast<float*>(ws);
}

// P = W1[:, 64:128] e + b1, Q = W1[:, 128:192] e, U = Wf[:, 64:128] e + bf, V = Wf[:, 128:192] e with
// e = W_init s + b_init (ipa_pytorch.py:233-241) are all linear in the node row s: compose them once into
// Wc [512, 128] (rows P 0-191 | Q 192-383 | U 384-447 | V 448-511) and bc [512], so the node-layer chain can emit
// the four terms straight from s as extra layers.
__global__ void edge_compose_terms_kernel(const float* __restrict__ w_init, const float* __restrict__ b_init,
                                          const float* __restrict__ w1, const float* __restrict__ b1,
                                          const float* __restrict__ wf, const float* __restrict__ bf,
                                          float* __restrict__ wc, float* __restrict__ bc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // (row n of Wc, column k) or, past 512 * 128, a bias entry
  const bool is_bias = idx >= 512 * 128;
  const int n = is_bias ? idx - 512 * 128 : idx >> 7, k = idx & 127;
  if (n >= 512) return;
  const float* row;                                           // 64 coefficients over e
  float add = 0.f;
  if (n < 192) { row = w1 + (size_t)n * 192 + 64; add = b1[n]; }
  else if (n < 384) { row = w1 + (size_t)(n - 192) * 192 + 128; }
  else if (n < 448) { row = wf + (size_t)(n - 384) * 192 + 64; add = bf[n - 384]; }
  else { row = wf + (size_t)(n - 448) * 192 + 128; }
  float acc = 0.f;
  if (is_bias) {
    for (int c = 0; c < 64; ++c) acc = fmaf(row[c], b_init[c], acc);
    bc[n] = acc + add;
  } else {
    for (int c = 0; c < 64; ++c) acc = fmaf(row[c], w_init[c * 128 + k], acc);
    wc[(size_t)n * 128 + k] = acc;
  }
}

int launch_edge_compose_terms(const float* w_init, const float* b_init, const float* w1, const float* b1,
                              const float* wf, const float* bf, float* wc, float* bc, cudaStream_t st) {
  edge_compose_terms_kernel<<<(512 * 128 + 512 + 255) / 256, 256, 0, st>>>(w_init, b_init, w1, b1, wf, bf, wc, bc);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_edge_transition(const float* s, const float* z_in, const float* w_init, const float* b_init,
                           const float* w1, const float* b1, const float* w2, const float* b2, const float* wf,
                           const float* bf, const float* ln_g, const float* ln_b, const float* mask, float* z_out,
                           void* workspace, size_t workspace_bytes, int B, int L, cudaStream_t st,
                           const void* prepacked_weights, bool terms_ready) {
  if (B == 0 || L == 0) return PF_OK;
  PF_REQUIRE(workspace_bytes >= edge_workspace_bytes(B, L), PF_ERR_WORKSPACE_TOO_SMALL);
  PF_REQUIRE(!terms_ready || opt_edge_impl() != 0, PF_ERR_BAD_CONFIG);
  const int M = B * L;
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  float* e = reinterpret_cast<float*>(ws); ws += align256((size_t)M * 64 * 4);
  if (!terms_ready) PF_TRY(launch_linear(s, w_init, b_init, nullptr, nullptr, e, M, 128, 64, 0, st));
  if (opt_edge_impl() == 0) {
    EdgeArgs a{e, z_in, w1, b1, w2, b2, wf, bf, ln_g, ln_b, mask, z_out, B, L};
    dim3 grid((L * L + E0_ROWS - 1) / E0_ROWS, B);
    profile_begin(1, st);
    edge_transition_v0_kernel<<<grid, 256, E0_SMEM, st>>>(a);
    profile_end(1, st);
    PF_CHECK_LAUNCH();
    return PF_OK;
  }
  float* P = reinterpret_cast<float*>(ws); ws += align256((size_t)M * 192 * 4);
  float* Q = reinterpret_cast<float*>(ws); ws += align256((size_t)M * 192 * 4);
  float* U = reinterpret_cast<float*>(ws); ws += align256((size_t)M * 64 * 4);
  float* V = reinterpret_cast<float*>(ws); ws += align256((size_t)M * 64 * 4);
  uint4* w2pack = reinterpret_cast<uint4*>(ws); ws += align256(E1_W2_BYTES);
  uint4* stream = reinterpret_cast<uint4*>(ws); ws += align256(E1_STREAM_BYTES);
  void* upack = ws;
  if (!terms_ready) {
    PF_TRY(launch_linear_ld(e, w1 + 64, 192, b1, P, M, 64, 192, st));
    PF_TRY(launch_linear_ld(e, w1 + 128, 192, nullptr, Q, M, 64, 192, st));
    PF_TRY(launch_linear_ld(e, wf + 64, 192, bf, U, M, 64, 64, st));
    PF_TRY(launch_linear_ld(e, wf + 128, 192, nullptr, V, M, 64, 64, st));
  }
  if (opt_edge_impl() == 2)
    return launch_edge_umma(z_in, P, Q, U, V, w1, w2, wf, b2, ln_g, ln_b, mask, z_out, upack, B, L, st,
                            prepacked_weights);
  {
    const int n = (E1_W2_BYTES + E1_STREAM_BYTES) / 16;
    edge_pack_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, w2, wf, w2pack, stream);
    PF_CHECK_LAUNCH();
  }
  Edge1Args a;
  a.z_in = z_in; a.P = P; a.Q = Q; a.U = U; a.V = V; a.b2 = b2; a.ln_g = ln_g; a.ln_b = ln_b; a.mask = mask;
  a.w2pack = w2pack; a.stream = stream; a.z_out = z_out; a.B = B; a.L = L;
  a.tiles_per_complex = (L * L + E1_ROWS - 1) / E1_ROWS;
  a.total_tiles = a.tiles_per_complex * B;
  const int grid = a.total_tiles < num_sms() ? a.total_tiles : num_sms();
  profile_begin(1, st);
  edge_transition_v1_kernel<<<grid, 256, E1_SMEM, st>>>(a);
  profile_end(1, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

}  // namespace pf

extern "C" {

size_t pf_edge_transition_workspace_bytes(int B, int L) { return pf::edge_workspace_bytes(B, L); }

int pf_edge_transition(const float* s, const float* z_in, const float* w_init, const float* b_init, const float* w1,
                       const float* b1, const float* w2, const float* b2, const float* wf, const float* bf,
                       const float* ln_g, const float* ln_b, const float* mask, float* z_out, void* workspace,
                       size_t workspace_bytes, int B, int L, void* stream) {
  PF_REQUIRE(s && z_in && w_init && b_init && w1 && b1 && w2 && b2 && wf && bf && ln_g && ln_b && mask && z_out &&
                 workspace, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  PF_REQUIRE(pf::aligned16(z_in) && pf::aligned16(z_out) && pf::aligned16(workspace), PF_ERR_MISALIGNED);
  return pf::launch_edge_transition(s, z_in, w_init, b_init, w1, b1, w2, b2, wf, bf, ln_g, ln_b, mask, z_out,
                                    workspace, workspace_bytes, B, L, pf::as_stream(stream));
}

}  // extern "C"
