// K8, variant 2: edge transition on the 5th-generation tensor cores (tcgen05.mma, accumulators and the
// A operand in tensor memory, CTA pairs).  Same arithmetic as variant 1 (pf_edge.cu): 3xFP16 split precision
// (hi*hi + lo*hi + hi*lo, fp32 accumulate) and the per-residue parts of W1 x and W_f x hoisted out of the
// pair loop (P_i + Q_j, U_i + V_j); reference: models_con/ipa_pytorch.py:233-248, mask models_con/ga.py:118.
//
// Work decomposition.  A cluster of two CTAs (one SM pair) owns a [16 i x 16 j] block of pairs of one complex;
// CTA r of the pair takes the 8 columns j0+8r .. j0+8r+7, i.e. 128 pair rows = the 128 lanes of its tensor
// memory (row = 8*i_local + j_local).  One tcgen05.mma.cta_group::2 covers both CTAs (M = 256).
//
// Shared memory (per CTA, 128 KB): this CTA's HALF of the N rows of every weight matrix as fp16 hi and lo
// parts in the K-major no-swizzle core-matrix layout (the pair together holds each matrix once):
//   B1 = [W1[:, 0:64] ; Wf[:, 0:64]]  N = 256, K = 64      B2 = W2  N = 192, K = 192      B3 = Wf  N = 64, K = 192
//
// Tensor memory (512 columns x 128 lanes per CTA):
//   [  0,192)  acc1 = z W1z^T            -> in place: h1 = relu(acc1 + P_i + Q_j) as packed fp16 hi | lo
//   [192,256)  acc3 = z Wfz^T (+= h2 Wf^T)
//   [256,448)  acc2 = h1 W2^T            -> in place: h2 = relu(acc2 + b2) as packed fp16 hi | lo
//   [448,512)  A0   = z as packed fp16 hi | lo
// A 32-column fp32 chunk turns into 16 columns of packed hi pairs followed by 16 columns of packed lo pairs
// at the same place, which is exactly the A operand (K = 32) of the next layer's MMAs - activations never
// touch shared memory or registers of another thread.
//
// Roles: warps 0-3 = 128 row threads (tensor-memory lane = pair row): stage z, run the three epilogues;
// warp 4 = tensor-memory allocation and, in the leader CTA, the single MMA-issuing thread.  Layer k+1's MMAs
// on K chunk c start as soon as the epilogue of layer k has converted chunk c (one mbarrier per chunk), so the
// tensor pipe runs underneath the epilogue.
#include "pf_common.cuh"
#include "pf_split.cuh"
#include "pf_umma.cuh"

namespace pf {

using namespace umma;

constexpr int EU_THREADS = 160;
constexpr int EU_RANK_BYTES = 131072;  // packed weights per CTA rank
constexpr int EU_B1_HI = 0, EU_B1_LO = 16384, EU_B2_HI = 32768, EU_B2_LO = 69632, EU_B3_HI = 106496, EU_B3_LO = 118784;
constexpr int EU_N1 = 256, EU_N2 = 192, EU_N3 = 64;          // full N of the three GEMMs
constexpr int EU_NL1 = 128, EU_NL2 = 96, EU_NL3 = 32;        // N rows held per CTA
constexpr uint32_t EU_COL_ACC1 = 0, EU_COL_ACC3 = 192, EU_COL_ACC2 = 256, EU_COL_A0 = 448;
// barriers
constexpr int EU_BAR_A0 = 0, EU_BAR_ACC1 = 1, EU_BAR_H1 = 2 /* +6 */, EU_BAR_ACC2 = 8, EU_BAR_H2 = 9 /* +6 */,
              EU_BAR_ACC3 = 15, EU_NBARS = 16;
constexpr int EU_SMEM_BARS = EU_RANK_BYTES;                       // 16 x 8 B
constexpr int EU_SMEM_TMEM_PTR = EU_SMEM_BARS + EU_NBARS * 8;     // 4 B (+12 pad)
constexpr int EU_SMEM_B2 = EU_SMEM_TMEM_PTR + 16;                 // 192 floats
constexpr int EU_SMEM_LNG = EU_SMEM_B2 + 192 * 4;                 // 64 floats
constexpr int EU_SMEM_LNB = EU_SMEM_LNG + 64 * 4;                 // 64 floats
constexpr int EU_SMEM_TOTAL = EU_SMEM_LNB + 64 * 4;

// Repack fp32 weights into the per-rank shared-memory image (fp16 hi / lo, 16-byte units = 8 consecutive k
// of one output row n, ordered [k/8][n_local]).
__global__ void edge_umma_pack_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                      const float* __restrict__ wf, uint4* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int UNITS = EU_RANK_BYTES / 16;  // 8192 per rank
  if (idx >= 2 * UNITS) return;
  const int rank = idx / UNITS;
  int u = idx % UNITS;
  // which matrix / part
  const float* src;
  bool lo;
  int nl, kc;
  if (u < 2048) {  // B1: hi 1024 units, lo 1024 units; unit = kc * 128 + nl
    lo = u >= 1024; u &= 1023;
    kc = u / EU_NL1; nl = u % EU_NL1;
    const int n = rank * EU_NL1 + nl;
    src = (n < 192 ? w1 + (size_t)n * 192 : wf + (size_t)(n - 192) * 192) + kc * 8;
  } else if (u < 2048 + 4608) {  // B2: 2304 + 2304; unit = kc * 96 + nl
    u -= 2048;
    lo = u >= 2304; u %= 2304;
    kc = u / EU_NL2; nl = u % EU_NL2;
    src = w2 + (size_t)(rank * EU_NL2 + nl) * 192 + kc * 8;
  } else {  // B3: 768 + 768; unit = kc * 32 + nl
    u -= 2048 + 4608;
    lo = u >= 768; u %= 768;
    kc = u / EU_NL3; nl = u % EU_NL3;
    src = wf + (size_t)(rank * EU_NL3 + nl) * 192 + kc * 8;
  }
  uint32_t hi[4], lw[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pair(src[2 * q], src[2 * q + 1], hi[q], lw[q]);
  out[idx] = lo ? make_uint4(lw[0], lw[1], lw[2], lw[3]) : make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

struct EdgeUArgs {
  const float* z_in;  // [B, L, L, 64]
  const float* P;     // [B*L, 192]  W1[:,64:128] e_i + b1
  const float* Q;     // [B*L, 192]  W1[:,128:192] e_j
  const float* U;     // [B*L, 64]   Wf[:,64:128] e_i + bf
  const float* V;     // [B*L, 64]   Wf[:,128:192] e_j
  const float* b2; const float* ln_g; const float* ln_b; const float* mask;
  const uint4* wpack;  // 2 x EU_RANK_BYTES
  float* z_out;
  int B, L, tiles_1d, total_blocks;  // tiles_1d = ceil(L / 16); total_blocks = B * tiles_1d^2
};

// 32 fp32 values -> 16 packed hi pairs + 16 packed lo pairs, stored over the chunk they came from.
__device__ __forceinline__ void store_split_chunk(uint32_t taddr, const float (&v)[32]) {
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) split_pair(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
  tmem_st16(taddr, hi);
  tmem_st16(taddr + 16, lo);
}

// 3xFP16 MMAs of one 32-wide K chunk (2 K steps) of an A operand living at tensor-memory column a_col.
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a_col, uint32_t b_hi, uint32_t b_lo, int kstep0,
                                            uint32_t nl, uint32_t idesc, bool first_overwrites) {
  const uint32_t kstep_bytes = 2 * nl * 16, lbo = nl * 16;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint64_t dh = smem_desc(b_hi + (kstep0 + s) * kstep_bytes, lbo, 128);
    const uint64_t dl = smem_desc(b_lo + (kstep0 + s) * kstep_bytes, lbo, 128);
    const uint32_t a_hi = a_col + 8 * s, a_lo = a_col + 16 + 8 * s;
    mma_pair_ts(d_tmem, a_lo, dh, idesc, (first_overwrites && s == 0) ? 0u : 1u);
    mma_pair_ts(d_tmem, a_hi, dl, idesc, 1u);
    mma_pair_ts(d_tmem, a_hi, dh, idesc, 1u);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EU_THREADS, 1) edge_transition_umma_kernel(EdgeUArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + EU_SMEM_BARS;
  auto bar = [&](int i) { return bars + 8u * i; };
  float* sB2 = reinterpret_cast<float*>(smem + EU_SMEM_B2);
  float* sG = reinterpret_cast<float*>(smem + EU_SMEM_LNG);
  float* sBt = reinterpret_cast<float*>(smem + EU_SMEM_LNB);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + EU_SMEM_TMEM_PTR);

  // ---- one-time setup ----------------------------------------------------------------------------
  if (warp == 4) tmem_alloc_pair(sbase + EU_SMEM_TMEM_PTR, 512);
  if (tid == 0) {
    // barriers the MMA thread waits on collect one arrival per row warp of both CTAs; the others one commit
    mbar_init(bar(EU_BAR_A0), 8);
    mbar_init(bar(EU_BAR_ACC1), 1);
    mbar_init(bar(EU_BAR_ACC2), 1);
    mbar_init(bar(EU_BAR_ACC3), 1);
    for (int c = 0; c < 6; ++c) { mbar_init(bar(EU_BAR_H1 + c), 8); mbar_init(bar(EU_BAR_H2 + c), 8); }
    fence_mbar_init();
  }
  {
    const uint4* src = a.wpack + (size_t)rank * (EU_RANK_BYTES / 16);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < EU_RANK_BYTES / 16; i += EU_THREADS) dst[i] = src[i];
    for (int i = tid; i < 192; i += EU_THREADS) sB2[i] = a.b2[i];
    if (tid < 64) { sG[tid] = a.ln_g[tid]; sBt[tid] = a.ln_b[tid]; }
  }
  fence_proxy_async_smem();   // weights were written through the generic proxy, the tensor core reads them async
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const int L = a.L, T1 = a.tiles_1d;
  const uint32_t first = cluster_id_x(), stride = num_clusters_x();

  if (warp < 4) {
    // ================================ row threads ================================================
    const int rl = warp * 32 + lane;                 // tensor-memory lane = pair row of this CTA's tile
    const uint32_t tlane = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t it = 0;
    for (uint32_t blk = first; blk < (uint32_t)a.total_blocks; blk += stride, ++it) {
      const uint32_t ph = it & 1u;
      const int b = blk / (T1 * T1), rem = blk % (T1 * T1);
      const int i = (rem / T1) * 16 + (rl >> 3);
      const int j = (rem % T1) * 16 + 8 * (int)rank + (rl & 7);
      const bool ok = i < L && j < L;
      const int ic = i < L ? i : L - 1, jc = j < L ? j : L - 1;
      const size_t rowb = (size_t)b * L;
      const size_t zoff = ((rowb + ic) * L + jc) * CZ;

      // ---- stage z as packed fp16 hi | lo (A0)
      {
        const float4* zp = reinterpret_cast<const float4*>(a.z_in + zoff);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 t4 = ok ? __ldg(zp + c * 8 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
          }
          store_split_chunk(tmem + tlane + EU_COL_A0 + 32 * c, v);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar(EU_BAR_A0), 0);
      }

      // ---- epilogue 1: h1 = relu(acc1 + P_i + Q_j)
      mbar_wait(bar(EU_BAR_ACC1), ph);
      tc_fence_after();
      {
        const float4* Pp = reinterpret_cast<const float4*>(a.P + (rowb + ic) * 192);
        const float4* Qp = reinterpret_cast<const float4*>(a.Q + (rowb + jc) * 192);
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem + tlane + EU_COL_ACC1 + 32 * c, r);
          float4 pq[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 p4 = __ldg(Pp + c * 8 + q), q4 = __ldg(Qp + c * 8 + q);
            pq[q] = make_float4(p4.x + q4.x, p4.y + q4.y, p4.z + q4.z, p4.w + q4.w);
          }
          tc_wait_ld();
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[4 * q] = fmaxf(__uint_as_float(r[4 * q]) + pq[q].x, 0.f);
            v[4 * q + 1] = fmaxf(__uint_as_float(r[4 * q + 1]) + pq[q].y, 0.f);
            v[4 * q + 2] = fmaxf(__uint_as_float(r[4 * q + 2]) + pq[q].z, 0.f);
            v[4 * q + 3] = fmaxf(__uint_as_float(r[4 * q + 3]) + pq[q].w, 0.f);
          }
          store_split_chunk(tmem + tlane + EU_COL_ACC1 + 32 * c, v);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(bar(EU_BAR_H1 + c), 0);
        }
      }

      // ---- epilogue 2: h2 = relu(acc2 + b2)
      mbar_wait(bar(EU_BAR_ACC2), ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem + tlane + EU_COL_ACC2 + 32 * c, r);
        tc_wait_ld();
        float v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = fmaxf(__uint_as_float(r[q]) + sB2[32 * c + q], 0.f);
        store_split_chunk(tmem + tlane + EU_COL_ACC2 + 32 * c, v);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar(EU_BAR_H2 + c), 0);
      }

      // ---- epilogue 3: y = acc3 + U_i + V_j; LayerNorm_64; pair mask; store
      mbar_wait(bar(EU_BAR_ACC3), ph);
      tc_fence_after();
      {
        float y[64];
        const float4* Up = reinterpret_cast<const float4*>(a.U + (rowb + ic) * 64);
        const float4* Vp = reinterpret_cast<const float4*>(a.V + (rowb + jc) * 64);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem + tlane + EU_COL_ACC3 + 32 * c, r);
          tc_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 u4 = __ldg(Up + c * 8 + q), v4 = __ldg(Vp + c * 8 + q);
            y[32 * c + 4 * q] = __uint_as_float(r[4 * q]) + (u4.x + v4.x);
            y[32 * c + 4 * q + 1] = __uint_as_float(r[4 * q + 1]) + (u4.y + v4.y);
            y[32 * c + 4 * q + 2] = __uint_as_float(r[4 * q + 2]) + (u4.z + v4.z);
            y[32 * c + 4 * q + 3] = __uint_as_float(r[4 * q + 3]) + (u4.w + v4.w);
          }
        }
        tc_fence_before();   // acc3 has been read: the next tile's MMAs may overwrite it after our next arrive
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 64; ++q) s += y[q];
        const float mu = s * (1.0f / 64.0f);
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < 64; ++q) { const float d = y[q] - mu; ss += d * d; }
        const float rstd = 1.0f / sqrtf(ss * (1.0f / 64.0f) + 1e-5f);
        if (ok) {
          const float pm = a.mask[rowb + i] * a.mask[rowb + j];
          float4* op = reinterpret_cast<float4*>(a.z_out + zoff);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float4 o;
            o.x = ((y[4 * q] - mu) * rstd * sG[4 * q] + sBt[4 * q]) * pm;
            o.y = ((y[4 * q + 1] - mu) * rstd * sG[4 * q + 1] + sBt[4 * q + 1]) * pm;
            o.z = ((y[4 * q + 2] - mu) * rstd * sG[4 * q + 2] + sBt[4 * q + 2]) * pm;
            o.w = ((y[4 * q + 3] - mu) * rstd * sG[4 * q + 3] + sBt[4 * q + 3]) * pm;
            op[q] = o;
          }
        }
      }
    }
  } else if (rank == 0 && lane == 0) {
    // ================================ MMA issuer (leader CTA, one thread) ==========================
    const uint32_t id1 = idesc_f16(256, EU_N1), id2 = idesc_f16(256, EU_N2), id3 = idesc_f16(256, EU_N3);
    uint32_t it = 0;
    for (uint32_t blk = first; blk < (uint32_t)a.total_blocks; blk += stride, ++it) {
      const uint32_t ph = it & 1u;
      // layer 1 (+ the z part of layer 3): [acc1 | acc3] = A0 [W1z ; Wfz]^T, K = 64
      mbar_wait(bar(EU_BAR_A0), ph);
      tc_fence_after();
      issue_chunk(tmem + EU_COL_ACC1, tmem + EU_COL_A0, sbase + EU_B1_HI, sbase + EU_B1_LO, 0, EU_NL1, id1, true);
      issue_chunk(tmem + EU_COL_ACC1, tmem + EU_COL_A0 + 32, sbase + EU_B1_HI, sbase + EU_B1_LO, 2, EU_NL1, id1, false);
      commit_pair(bar(EU_BAR_ACC1));
      // layer 2: acc2 = h1 W2^T, K chunk by K chunk as epilogue 1 produces h1
      for (int c = 0; c < 6; ++c) {
        mbar_wait(bar(EU_BAR_H1 + c), ph);
        tc_fence_after();
        issue_chunk(tmem + EU_COL_ACC2, tmem + EU_COL_ACC1 + 32 * c, sbase + EU_B2_HI, sbase + EU_B2_LO, 2 * c, EU_NL2,
                    id2, c == 0);
      }
      commit_pair(bar(EU_BAR_ACC2));
      // layer 3: acc3 += h2 Wf^T
      for (int c = 0; c < 6; ++c) {
        mbar_wait(bar(EU_BAR_H2 + c), ph);
        tc_fence_after();
        issue_chunk(tmem + EU_COL_ACC3, tmem + EU_COL_ACC2 + 32 * c, sbase + EU_B3_HI, sbase + EU_B3_LO, 2 * c, EU_NL3,
                    id3, false);
      }
      commit_pair(bar(EU_BAR_ACC3));
    }
  }

  // ---- teardown: nobody leaves (or frees tensor memory) while the pair may still touch this CTA
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 4) tmem_dealloc_pair(tmem, 512);
}

void edge_umma_init() {
  cudaFuncSetAttribute(edge_transition_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, EU_SMEM_TOTAL);
}

size_t edge_umma_pack_bytes() { return 2 * (size_t)EU_RANK_BYTES; }

int launch_edge_umma(const float* z_in, const float* P, const float* Q, const float* U, const float* V,
                     const float* w1, const float* w2, const float* wf, const float* b2, const float* ln_g,
                     const float* ln_b, const float* mask, float* z_out, void* wpack, int B, int L, cudaStream_t st) {
  {
    const int n = 2 * EU_RANK_BYTES / 16;
    edge_umma_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, w2, wf, static_cast<uint4*>(wpack));
    PF_CHECK_LAUNCH();
  }
  EdgeUArgs a;
  a.z_in = z_in; a.P = P; a.Q = Q; a.U = U; a.V = V; a.b2 = b2; a.ln_g = ln_g; a.ln_b = ln_b; a.mask = mask;
  a.wpack = static_cast<const uint4*>(wpack); a.z_out = z_out; a.B = B; a.L = L;
  a.tiles_1d = (L + 15) / 16;
  a.total_blocks = B * a.tiles_1d * a.tiles_1d;
  int clusters = num_sms() / 2;
  if (clusters > a.total_blocks) clusters = a.total_blocks;
  if (clusters < 1) clusters = 1;
  profile_begin(1, st);
  edge_transition_umma_kernel<<<2 * clusters, EU_THREADS, EU_SMEM_TOTAL, st>>>(a);
  profile_end(1, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

}  // namespace pf
