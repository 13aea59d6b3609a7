// K8, variant 2: edge transition on the 5th-generation tensor cores (tcgen05.mma, accumulators and the
// A operand in tensor memory, CTA pairs).  Same arithmetic as variant 1 (pf_edge.cu): 3xFP16 split precision
// (hi*hi + lo*hi + hi*lo, fp32 accumulate) and the per-residue parts of W1 x and W_f x hoisted out of the
// pair loop (P_i + Q_j, U_i + V_j); reference: models_con/ipa_pytorch.py:233-248, mask models_con/ga.py:118.
//
// Work decomposition.  A cluster of two CTAs (one SM pair) owns a [16 i x 16 j] block of pairs of one complex;
// CTA r of the pair takes the 8 columns j0+8r .. j0+8r+7, i.e. 128 pair rows = the 128 lanes of its tensor
// memory (row = 8*i_local + j_local).  One tcgen05.mma.cta_group::2 covers both CTAs (M = 256).
//
// Shared memory (per CTA):
//   weights, 128 KB: this CTA's HALF of the N rows of every weight matrix as fp16 hi and lo parts in the
//     K-major no-swizzle core-matrix layout (the pair together holds each matrix once):
//       B1a = W1[:, 0:64] (N 192, K 64)  B1b = Wf[:, 0:64] (N 64, K 64)  B2 = W2 (N 192, K 192)  B3 = Wf (N 64, K 192)
//   selector S, 8 KB: constant 0/1 matrix [128 rows x 32] with S[row, row/8] = S[row, 16 + j_local(row)] = 1.
//     S x [P rows i0..i0+15 ; Q rows j0..j0+15] is P_i + Q_j for every pair row - the broadcast add of the
//     hoisted per-residue terms runs on the tensor core instead of 128 gathered loads per row.
//   PQ / UV tiles, 12 + 4 KB: those 32 residue rows (fp16 hi, lo) for the block, bulk-copied per tile from a
//     pre-packed image (MN-major core matrices), B operand of the selector MMAs.
//   z tiles, 2 x 32 KB: 32 TMA boxes of [8 j x 32 channels] (1 KB, 128-byte swizzle, so that the 16-byte
//     row-per-thread accesses are conflict-free) loaded through a 3-D tensor map over z [B*L, L, 64]; the output
//     tile is staged in the same boxes and stored through a second map (TMA clips the ragged j edge).
//
// Tensor memory (512 columns x 128 lanes per CTA):
//   [  0,192)  acc1 = S [P;Q] + z W1z^T   -> in place: h1 = relu(acc1) as packed fp16 hi | lo
//   [192,256)  acc3 = S [U;V] + z Wfz^T (+= h2 Wf^T)
//   [256,448)  acc2 = h1 W2^T             -> in place: h2 = relu(acc2 + b2) as packed fp16 hi | lo
//   [448,512)  A0   = z as packed fp16 hi | lo
// A 32-column fp32 chunk turns into 16 columns of packed hi pairs followed by 16 columns of packed lo pairs
// at the same place, which is exactly the A operand (K = 32) of the next layer's MMAs - activations never
// touch shared memory or registers of another thread.
//
// Roles: warps 0-7 = row threads (tensor-memory lane = pair row; warps w and w+4 split the columns of a row);
// warp 8 = tensor-memory allocation and, in the leader CTA, the MMA issuer; warp 9 = bulk-copy producer of the
// PQ / UV tiles.  Layer k+1's MMAs on K chunk c start as soon as the epilogue of layer k has converted chunk c
// (one mbarrier per chunk), and the next tile's first layer is issued before this tile's LayerNorm epilogue,
// so the tensor pipe runs underneath the epilogues.
#include <cuda.h>   // CUtensorMap (the encoder is fetched at run time through cudaGetDriverEntryPoint)

#include "pf_common.cuh"
#include "pf_split.cuh"
#include "pf_umma.cuh"

namespace pf {

using namespace umma;

constexpr int EU_THREADS = 320;        // 8 row warps + MMA / allocation warp + producer warp
constexpr int EU_RANK_BYTES = 131072;  // packed weights per CTA rank
// byte offsets of the weight parts inside a rank image; NL = N rows held per CTA
// (W2 is split by output columns into n 0..127 and n 128..191 so that the epilogue of the first part runs
// while the second part is still on the tensor core)
constexpr int EU_NL1A = 96, EU_NL1B = 32, EU_NL2A = 64, EU_NL2B = 32, EU_NL3 = 32;
constexpr int EU_B1A_HI = 0, EU_B1A_LO = 12288, EU_B1B_HI = 24576, EU_B1B_LO = 28672, EU_B2A_HI = 32768,
              EU_B2A_LO = 57344, EU_B2B_HI = 81920, EU_B2B_LO = 94208, EU_B3_HI = 106496, EU_B3_LO = 118784;
constexpr uint32_t EU_COL_ACC1 = 0, EU_COL_ACC3 = 192, EU_COL_ACC2 = 256, EU_COL_ACC2B = 384, EU_COL_A0 = 448;
// shared-memory map
constexpr int EU_SMEM_SEL = EU_RANK_BYTES;            // 8192: [k/8 (4)][row (128)][16 B]
constexpr int EU_SMEM_PQ = EU_SMEM_SEL + 8192;        // 12288: P_hi | Q_hi | P_lo | Q_lo, each [kg 2][ng 12][128 B]
constexpr int EU_SMEM_UV = EU_SMEM_PQ + 12288;        // 4096:  U_hi | V_hi | U_lo | V_lo, each [kg 2][ng 4][128 B]
constexpr int EU_ZBUF = 32768;                        // [half 2][i_local 16] boxes of 8 rows x 128 B
constexpr int EU_SMEM_Z = EU_SMEM_UV + 4096;          // 2 buffers (1024-byte aligned: 128B-swizzled boxes)
constexpr int EU_SMEM_RED = EU_SMEM_Z + 2 * EU_ZBUF;  // LayerNorm exchange: [tile parity 2][grp 2][mean, M2][128 rows] floats
constexpr int EU_SMEM_BARS = EU_SMEM_RED + 4096;
static_assert(EU_SMEM_Z % 1024 == 0, "swizzled TMA boxes need 1024-byte alignment");
// barriers
enum {
  EU_BAR_A0 = 0,    // leader; 16 row warps: A0 of the tile staged
  EU_BAR_ACC1,      // both;   commit: acc1 complete (P/Q tile and acc1's previous content consumed)
  EU_BAR_H1,        // leader; +6, 8 row warps each: h1 chunk c in place
  EU_BAR_ACC2A = EU_BAR_H1 + 6,  // both; commit: acc2 columns 0..127 complete
  EU_BAR_ACC2B,     // both;   commit: acc2 columns 128..191 complete
  EU_BAR_PB,        // both;   commit: the acc3 = S[U;V] + z Wfz part is done: A0 and the U/V tile are free
  EU_BAR_H2,        // leader; +6
  EU_BAR_ACC3 = EU_BAR_H2 + 6,   // both; commit
  EU_BAR_ACC3F,     // leader; 16 row warps: acc3 has been read
  EU_BAR_PQF,       // local;  P/Q tile landed (transaction bytes)
  EU_BAR_UVF,       // local;  U/V tile landed
  EU_BAR_PQR,       // leader; 2 producers: P/Q tile landed in both CTAs
  EU_BAR_UVR,       // leader; 2 producers
  EU_BAR_Z,         // local;  +2, 256 row threads + transaction bytes: z tile landed in buffer b
  EU_NBARS = EU_BAR_Z + 2
};
constexpr int EU_SMEM_TMEM_PTR = EU_SMEM_BARS + ((EU_NBARS * 8 + 15) & ~15);
// b2 as B operands of one selector K step (every one of the 16 k rows = b2, so S[:, 0:16] x tile = b2 per row):
// part a (n 0..127, 64 per CTA): hi [kg 2][ng 8][8][16 B] = 2 KB | lo 2 KB; part b (n 128..191, 32 per CTA): 1 KB | 1 KB
constexpr int EU_SMEM_B2 = EU_SMEM_TMEM_PTR + 16;     // 6144 bytes
constexpr int EU_SMEM_LNG = EU_SMEM_B2 + 6144;        // 64 floats
constexpr int EU_SMEM_LNB = EU_SMEM_LNG + 64 * 4;     // 64 floats
constexpr int EU_SMEM_TOTAL = EU_SMEM_LNB + 64 * 4;
static_assert(EU_SMEM_TOTAL <= 232448, "shared memory budget");

// ---- pre-pack kernels ------------------------------------------------------------------------------
// fp32 weights -> per-rank shared-memory image (fp16 hi / lo; 16-byte unit = 8 consecutive k of one output
// row n; units ordered [k/8][n_local]).
__global__ void edge_umma_pack_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                              const float* __restrict__ wf, uint4* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int UNITS = EU_RANK_BYTES / 16;  // 8192 per rank
  if (idx >= 2 * UNITS) return;
  const int rank = idx / UNITS;
  int u = idx % UNITS;
  const float* src;
  bool lo;
  if (u < 1536) {            // B1a: 768 hi + 768 lo; unit = kc * 96 + nl
    lo = u >= 768; u %= 768;
    src = w1 + (size_t)(rank * EU_NL1A + u % EU_NL1A) * 192 + (u / EU_NL1A) * 8;
  } else if (u < 2048) {     // B1b: 256 + 256; unit = kc * 32 + nl
    u -= 1536;
    lo = u >= 256; u %= 256;
    src = wf + (size_t)(rank * EU_NL1B + u % EU_NL1B) * 192 + (u / EU_NL1B) * 8;
  } else if (u < 2048 + 3072) {  // B2a (n 0..127): 1536 + 1536; unit = kc * 64 + nl
    u -= 2048;
    lo = u >= 1536; u %= 1536;
    src = w2 + (size_t)(rank * EU_NL2A + u % EU_NL2A) * 192 + (u / EU_NL2A) * 8;
  } else if (u < 2048 + 4608) {  // B2b (n 128..191): 768 + 768; unit = kc * 32 + nl
    u -= 2048 + 3072;
    lo = u >= 768; u %= 768;
    src = w2 + (size_t)(128 + rank * EU_NL2B + u % EU_NL2B) * 192 + (u / EU_NL2B) * 8;
  } else {                   // B3: 768 + 768; unit = kc * 32 + nl
    u -= 2048 + 4608;
    lo = u >= 768; u %= 768;
    src = wf + (size_t)(rank * EU_NL3 + u % EU_NL3) * 192 + (u / EU_NL3) * 8;
  }
  uint32_t hi[4], lw[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pair(src[2 * q], src[2 * q + 1], hi[q], lw[q]);
  out[idx] = lo ? make_uint4(lw[0], lw[1], lw[2], lw[3]) : make_uint4(hi[0], hi[1], hi[2], hi[3]);
}

// Packed image of the hoisted per-residue terms (B operand of the selector MMAs, MN-major core matrices:
// 16-byte unit = 8 consecutive n of one residue row; 8 residue rows = one 128-byte core matrix).
//   wide parts (P, Q; 192 columns, 96 per CTA): [b][half][kg][ng 12][kr 8][16 B]
//   out  parts (U, V;  64 columns, 32 per CTA): [b][half][kg][ng  4][kr 8][16 B]
// kg = residue / 8 over Lp = 16*ceil(L/16) padded rows (zeros).  Region order: P_hi P_lo Q_hi Q_lo U_hi U_lo V_hi V_lo.
struct PackLayout {
  size_t wide_bytes, out_bytes;  // per region
  int KG;
  __host__ __device__ size_t region(int which) const {  // which: 0 P,1 Q (wide) 2 U,3 V (out); returns hi offset
    return which < 2 ? (size_t)which * 2 * wide_bytes : 4 * wide_bytes + (size_t)(which - 2) * 2 * out_bytes;
  }
};
static PackLayout pack_layout(int B, int L) {
  PackLayout p;
  p.KG = 2 * ((L + 15) / 16);
  p.wide_bytes = (size_t)B * 2 * p.KG * 1536;
  p.out_bytes = (size_t)B * 2 * p.KG * 512;
  return p;
}

__global__ void edge_umma_pack_terms_kernel(const float* __restrict__ P, const float* __restrict__ Q,
                                            const float* __restrict__ U, const float* __restrict__ V,
                                            unsigned char* __restrict__ out, PackLayout lay, int B, int L) {
  // one thread per (pair PU/QV, b, half, kg, ng16, kr)
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per_src = (size_t)B * 2 * lay.KG * 16 * 8;
  if (idx >= 2 * per_src) return;
  const int which = idx >= per_src;   // 0: P/U   1: Q/V
  size_t u = idx % per_src;
  const int kr = u % 8; u /= 8;
  const int ng = u % 16; u /= 16;
  const int kg = u % lay.KG; u /= lay.KG;
  const int half = u % 2;
  const int b = (int)(u / 2);
  const int res = kg * 8 + kr;
  const bool wide = ng < 12;
  float v[8];
  if (res < L) {
    const float* src = wide ? (which ? Q : P) + ((size_t)b * L + res) * 192 + half * 96 + ng * 8
                            : (which ? V : U) + ((size_t)b * L + res) * 64 + half * 32 + (ng - 12) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = src[q];
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = 0.f;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) split_pair(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
  const size_t slab = (size_t)(b * 2 + half) * lay.KG + kg;
  unsigned char* base;
  size_t off, lo_off;
  if (wide) { base = out + lay.region(which); off = (slab * 12 + ng) * 128 + kr * 16; lo_off = lay.wide_bytes; }
  else { base = out + lay.region(2 + which); off = (slab * 4 + (ng - 12)) * 128 + kr * 16; lo_off = lay.out_bytes; }
  *reinterpret_cast<uint4*>(base + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(base + lo_off + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct alignas(64) EdgeUArgs {
  CUtensorMap tm_in, tm_out;    // z_in / z_out as [B*L, L, 64] fp32, box [1, 8, 32], 128-byte swizzle
  const float* z_in;            // [B, L, L, 64]
  const unsigned char* terms;   // packed P/Q/U/V image (PackLayout)
  PackLayout lay;
  const float* b2; const float* ln_g; const float* ln_b; const float* mask;
  const uint4* wpack;           // 2 x EU_RANK_BYTES
  float* z_out;
  long long* dbg;
  int B, L, tiles_1d, total_blocks;  // tiles_1d = ceil(L / 16); total_blocks = B * tiles_1d^2
  int drop_terms;               // "edge_terms" option: bit g set = GEMM g (z W1z, z Wfz, h1 W2, h2 Wf) drops A_hi W_lo
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
// drop_hl: leave out the A_hi x W_lo product (two passes; "edge_terms" option, profiles/r2_edge_error_budget.txt)
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a_col, uint32_t b_hi, uint32_t b_lo, int kstep0,
                                            uint32_t nl, uint32_t idesc, bool first_overwrites, bool drop_hl = false) {
  const uint32_t kstep_bytes = 2 * nl * 16, lbo = nl * 16;
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint64_t dh = smem_desc(b_hi + (kstep0 + s) * kstep_bytes, lbo, 128);
    const uint64_t dl = smem_desc(b_lo + (kstep0 + s) * kstep_bytes, lbo, 128);
    const uint32_t a_hi = a_col + 8 * s, a_lo = a_col + 16 + 8 * s;
    if (elect_one()) {
      mma_pair_ts(d_tmem, a_lo, dh, idesc, (first_overwrites && s == 0) ? 0u : 1u);
      if (!drop_hl) mma_pair_ts(d_tmem, a_hi, dl, idesc, 1u);
      mma_pair_ts(d_tmem, a_hi, dh, idesc, 1u);
    }
    __syncwarp();
  }
}
// D = S x [rows of the i block ; rows of the j block] (hi + lo): overwrites D.
//   piece layout in shared memory: X_hi | Y_hi | X_lo | Y_lo, each [kg 2][ng][128 B]  (X = P or U, Y = Q or V)
__device__ __forceinline__ void issue_selector(uint32_t d_tmem, uint32_t sel, uint32_t tile, uint32_t ng, uint32_t idesc) {
  const uint32_t piece = 2 * ng * 128, lbo = ng * 128;
#pragma unroll
  for (int s = 0; s < 2; ++s) {       // K step 0: the i rows (X), K step 1: the j rows (Y)
    const uint64_t da = smem_desc(sel + s * 4096, 2048, 128);
    const uint64_t bh = smem_desc(tile + s * piece, lbo, 128);
    const uint64_t bl = smem_desc(tile + (2 + s) * piece, lbo, 128);
    if (elect_one()) {
      mma_pair_ss(d_tmem, da, bh, idesc, s == 0 ? 0u : 1u);
      mma_pair_ss(d_tmem, da, bl, idesc, 1u);
    }
    __syncwarp();
  }
}

// D = S[:, 0:16] x (16 identical bias rows), hi then lo: D = bias for every row.  tile: hi | lo, each [kg 2][ng][128 B].
__device__ __forceinline__ void issue_bias(uint32_t d_tmem, uint32_t sel, uint32_t tile, uint32_t ng, uint32_t idesc) {
  const uint32_t piece = 2 * ng * 128, lbo = ng * 128;
  const uint64_t da = smem_desc(sel, 2048, 128);
  const uint64_t bh = smem_desc(tile, lbo, 128);
  const uint64_t bl = smem_desc(tile + piece, lbo, 128);
  if (elect_one()) {
    mma_pair_ss(d_tmem, da, bh, idesc, 0u);
    mma_pair_ss(d_tmem, da, bl, idesc, 1u);
  }
  __syncwarp();
}

// DBG: cluster 0 / CTA 0 stamps clock64() at every hand-off of its first 4 tiles into a.dbg
// ([3 actors: row group 0, row group 1, MMA warp][4 tiles][16 events]) - see scripts/gpu_edge_timeline.py.
template <bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(EU_THREADS, 1)
edge_transition_umma_kernel(const __grid_constant__ EdgeUArgs a) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + EU_SMEM_BARS;
  auto bar = [&](int i) { return bars + 8u * i; };
  float* sG = reinterpret_cast<float*>(__builtin_assume_aligned(smem + EU_SMEM_LNG, 16));
  float* sBt = reinterpret_cast<float*>(__builtin_assume_aligned(smem + EU_SMEM_LNB, 16));
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + EU_SMEM_TMEM_PTR);
  volatile float* sRed = reinterpret_cast<volatile float*>(smem + EU_SMEM_RED);

  // ---- one-time setup ----------------------------------------------------------------------------
  if ((sbase & 1023u) != 0) __trap();   // the swizzled TMA boxes assume a 1024-byte aligned window
  if (warp == 8) tmem_alloc_pair(sbase + EU_SMEM_TMEM_PTR, 512);
  if (tid == 32) { tma_prefetch_desc(&a.tm_in); tma_prefetch_desc(&a.tm_out); }
  if (tid == 0) {
    mbar_init(bar(EU_BAR_A0), 16);
    mbar_init(bar(EU_BAR_ACC1), 1);
    mbar_init(bar(EU_BAR_ACC2A), 1);
    mbar_init(bar(EU_BAR_ACC2B), 1);
    mbar_init(bar(EU_BAR_PB), 1);
    mbar_init(bar(EU_BAR_ACC3), 1);
    mbar_init(bar(EU_BAR_ACC3F), 16);
    mbar_init(bar(EU_BAR_PQF), 1);
    mbar_init(bar(EU_BAR_UVF), 1);
    mbar_init(bar(EU_BAR_PQR), 2);
    mbar_init(bar(EU_BAR_UVR), 2);
    mbar_init(bar(EU_BAR_Z), 256);
    mbar_init(bar(EU_BAR_Z + 1), 256);
    for (int c = 0; c < 6; ++c) { mbar_init(bar(EU_BAR_H1 + c), 8); mbar_init(bar(EU_BAR_H2 + c), 8); }
    fence_mbar_init();
  }
  {
    const uint4* src = a.wpack + (size_t)rank * (EU_RANK_BYTES / 16);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int i = tid; i < EU_RANK_BYTES / 16; i += EU_THREADS) dst[i] = src[i];
    // selector: unit (kc, row) holds k = 8 kc .. 8 kc + 7 of that row
    uint4* sel = reinterpret_cast<uint4*>(smem + EU_SMEM_SEL);
    for (int i = tid; i < 512; i += EU_THREADS) {
      const int kc = i >> 7, row = i & 127;
      const int k1 = row >> 3, k2 = 16 + 8 * (int)rank + (row & 7);
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int k = kc * 8 + q;
        if (k == k1 || k == k2) w[q >> 1] |= 0x3C00u << (16 * (q & 1));   // fp16 1.0
      }
      sel[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    // b2 tiles: unit u = ((part, hi|lo), kg, ng, kr): 8 consecutive n of b2, identical for every k row
    for (int i = tid; i < 384; i += EU_THREADS) {
      int u = i, n0, ngs, off;
      if (u < 256) { n0 = 64 * (int)rank; ngs = 8; off = 0; }            // part a: hi 128 units | lo 128 units
      else { u -= 256; n0 = 128 + 32 * (int)rank; ngs = 4; off = 4096; } // part b: hi 64 | lo 64
      const int per = 2 * ngs * 8;
      const bool lo = u >= per;
      const int ng = ((u % per) / 8) % ngs;
      uint32_t hi4[4], lo4[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) split_pair(a.b2[n0 + ng * 8 + 2 * q], a.b2[n0 + ng * 8 + 2 * q + 1], hi4[q], lo4[q]);
      reinterpret_cast<uint4*>(smem + EU_SMEM_B2 + off)[u] =
          lo ? make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]) : make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
    }
    if (tid < 64) { sG[tid] = a.ln_g[tid]; sBt[tid] = a.ln_b[tid]; }
  }
  fence_proxy_async_smem();   // generic-proxy writes above are read by the tensor core through the async proxy
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;

  const int L = a.L, T1 = a.tiles_1d;
  const uint32_t first = cluster_id_x(), stride = num_clusters_x(), nblk = (uint32_t)a.total_blocks;
  const bool dbg_on = DBG && first == 0 && rank == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 8);
  const int dbg_actor = warp == 8 ? 2 : (warp >> 2);
  auto stamp = [&](uint32_t it, int ev) {
    if (DBG && dbg_on && it < 4) a.dbg[(dbg_actor * 4 + it) * 16 + ev] = clock64();
  };

  if (warp < 8) {
    // ================================ row threads ================================================
    // Warps w and w+4 share the tensor-memory lanes 32*(w%4)..+31 (= pair rows) and split the columns:
    // group g = w/4 takes the 32-column chunks c = g, g+2, g+4 of each layer and channels 32g..32g+31 of z.
    const int grp = warp >> 2, quad = warp & 3;
    const int rl = quad * 32 + lane;                 // tensor-memory lane = pair row of this CTA's tile
    const uint32_t tlane = static_cast<uint32_t>(quad * 32) << 16;
    // z box of this thread's 8-row group and channel half: [8 rows][128 B], 16-byte chunk q of row r at q ^ r
    const int r8 = rl & 7;
    const uint32_t my_box = sbase + EU_SMEM_Z + (grp * 16 + (rl >> 3)) * 1024;      // + buffer * EU_ZBUF
    unsigned char* my_rowp = smem + EU_SMEM_Z + (grp * 16 + (rl >> 3)) * 1024 + r8 * 128;

    auto coords = [&](uint32_t blk, int& b, int& i, int& j) {
      b = blk / (T1 * T1);
      const int rem = blk % (T1 * T1);
      i = (rem / T1) * 16 + (rl >> 3);
      j = (rem % T1) * 16 + 8 * (int)rank + r8;
    };
    auto row_valid = [&](uint32_t blk) -> bool {
      if (blk >= nblk) return false;
      int b, i, j;
      coords(blk, b, i, j);
      return i < L && j < L;
    };
    // One thread per box (the first of its 8 rows) moves it; boxes entirely outside the complex are skipped,
    // partially outside ones are zero-filled / clipped by the TMA unit.
    auto request_z = [&](uint32_t blk, uint32_t buf) {
      int b, i, j;
      coords(blk, b, i, j);
      if (r8 == 0 && blk < nblk && i < L && j < L) {
        mbar_arrive_expect_tx(bar(EU_BAR_Z + buf), 1024);
        tma_load_3d(my_box + buf * EU_ZBUF, &a.tm_in, 32 * grp, j, b * L + i, bar(EU_BAR_Z + buf));
      } else {
        mbar_arrive(bar(EU_BAR_Z + buf));
      }
    };
    auto stage_a0 = [&](uint32_t blk, uint32_t n) {   // A0 of tile n from z buffer n & 1
      const uint32_t buf = n & 1u;
      mbar_wait(bar(EU_BAR_Z + buf), (n >> 1) & 1u);
      float v[32];
      if (row_valid(blk)) {
        const unsigned char* zp = my_rowp + buf * EU_ZBUF;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t4 = *reinterpret_cast<const float4*>(zp + ((q ^ r8) << 4));
          v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
        }
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      }
      store_split_chunk(tmem + tlane + EU_COL_A0 + 32 * grp, v);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(EU_BAR_A0), 0);
    };

    // pair mask of this thread's row for block blk (0 outside the complex)
    auto pair_mask = [&](uint32_t blk) -> float {
      if (!row_valid(blk)) return 0.f;
      int b, i, j;
      coords(blk, b, i, j);
      return __ldg(a.mask + (size_t)b * L + i) * __ldg(a.mask + (size_t)b * L + j);
    };
    // epilogue 3 of tile n (block blk): y = acc3; LayerNorm_64 (two groups x 32 channels); pair mask; the output
    // half row is staged in z buffer (n+1)&1 (its z was consumed when A0 of tile n+1 was staged) and bulk-stored.
    auto epilogue3 = [&](uint32_t blk, uint32_t n, float pm) {
      const uint32_t obuf = (n + 1) & 1u;
      mbar_wait(bar(EU_BAR_ACC3), n & 1u);
      tc_fence_after();
      float y[32];
      {
        uint32_t r[32];
        tmem_ld32(tmem + tlane + EU_COL_ACC3 + 32 * grp, r);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar(EU_BAR_ACC3F), 0);   // acc3 may be overwritten
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = __uint_as_float(r[q]);
      }
      // LayerNorm statistics: each group reduces its own 32 channels (mean, M2 about that mean); the two halves
      // are merged with the pairwise update of Chan et al.; only the two warps that share the rows synchronise.
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 32; q += 4) s += (y[q] + y[q + 1]) + (y[q + 2] + y[q + 3]);
      const float mean_g = s * (1.0f / 32.0f);
      float m2_g = 0.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) { const float d = y[q] - mean_g; m2_g += d * d; }
      volatile float* red = sRed + (n & 1u) * 512;
      red[(grp * 2 + 0) * 128 + rl] = mean_g;
      red[(grp * 2 + 1) * 128 + rl] = m2_g;
      asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quad) : "memory");
      const float mean_o = red[((grp ^ 1) * 2 + 0) * 128 + rl], m2_o = red[((grp ^ 1) * 2 + 1) * 128 + rl];
      const float mu = 0.5f * (mean_g + mean_o);
      const float dm = mean_o - mean_g;
      const float rstd = 1.0f / sqrtf((m2_g + m2_o + dm * dm * 16.0f) * (1.0f / 64.0f) + 1e-5f);
      {
        unsigned char* op = my_rowp + obuf * EU_ZBUF;
        const float* gG = sG + 32 * grp;
        const float* gB = sBt + 32 * grp;
        const float sc = rstd * pm;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o;
          o.x = (y[4 * q] - mu) * sc * gG[4 * q] + gB[4 * q] * pm;
          o.y = (y[4 * q + 1] - mu) * sc * gG[4 * q + 1] + gB[4 * q + 1] * pm;
          o.z = (y[4 * q + 2] - mu) * sc * gG[4 * q + 2] + gB[4 * q + 2] * pm;
          o.w = (y[4 * q + 3] - mu) * sc * gG[4 * q + 3] + gB[4 * q + 3] * pm;
          *reinterpret_cast<float4*>(op + ((q ^ r8) << 4)) = o;
        }
      }
      fence_proxy_async_smem();                 // generic-proxy stores -> visible to the TMA unit
      __syncwarp();                             // the 8 rows of a box live in 8 consecutive lanes
      int b, i, j;
      coords(blk, b, i, j);
      if (r8 == 0 && i < L && j < L) {
        tma_store_3d(&a.tm_out, 32 * grp, j, b * L + i, my_box + obuf * EU_ZBUF);
        bulk_commit();
      }
    };

    request_z(first, 0);
    request_z(first + stride, 1);
    if (first < nblk) stage_a0(first, 0);

    uint32_t it = 0;
    float pm_prev = 0.f;
    uint32_t blk_prev = 0;
    // one 32-column chunk of a hidden layer: fp32 accumulator -> relu(. + bias) -> packed fp16 hi | lo in place
    auto convert_chunk = [&](uint32_t col, const uint32_t (&r)[32], const float* bias, uint32_t done_bar) {
      float v[32];
      if (bias) {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = fmaxf(__uint_as_float(r[q]) + bias[q], 0.f);
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = fmaxf(__uint_as_float(r[q]), 0.f);
      }
      store_split_chunk(tmem + tlane + col, v);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(done_bar, 0);
    };
    for (uint32_t blk = first; blk < nblk; blk += stride, ++it) {
      const uint32_t ph = it & 1u;
      stamp(it, 0);
      const float pm_cur = pair_mask(blk);
      // ---- epilogue 3 of the previous tile, underneath this tile's first layer
      if (it > 0) epilogue3(blk_prev, it - 1, pm_prev);
      stamp(it, 11);
      // ---- epilogue 1: h1 = relu(acc1)   (P_i + Q_j + b1 already inside, via the selector MMAs)
      mbar_wait(bar(EU_BAR_ACC1), ph);
      tc_fence_after();
      stamp(it, 2);
      {
        uint32_t r[32];
        tmem_ld32(tmem + tlane + EU_COL_ACC1 + 32 * grp, r);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int c = grp + 2 * k;
          tc_wait_ld();
          uint32_t rc[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) rc[q] = r[q];
          if (k < 2) tmem_ld32(tmem + tlane + EU_COL_ACC1 + 32 * (c + 2), r);
          convert_chunk(EU_COL_ACC1 + 32 * c, rc, nullptr, bar(EU_BAR_H1 + c));
          stamp(it, 3 + k);
        }
      }

      // ---- next tile's A0 while layer 2 runs: A0 is free once the acc3 part of this tile has consumed it
      mbar_wait(bar(EU_BAR_PB), ph);
      if (blk + stride < nblk) stage_a0(blk + stride, it + 1);
      stamp(it, 1);

      // ---- epilogue 2: h2 = relu(acc2) (b2 was added by a selector MMA); columns 0..127 while columns 128..191
      //      are still being computed
      mbar_wait(bar(EU_BAR_ACC2A), ph);
      tc_fence_after();
      stamp(it, 6);
      {
        uint32_t r[32];
        tmem_ld32(tmem + tlane + EU_COL_ACC2 + 32 * grp, r);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int c = grp + 2 * k;
          tc_wait_ld();
          uint32_t rc[32];
#pragma unroll
          for (int q = 0; q < 32; ++q) rc[q] = r[q];
          if (k < 1) tmem_ld32(tmem + tlane + EU_COL_ACC2 + 32 * (c + 2), r);
          convert_chunk(EU_COL_ACC2 + 32 * c, rc, nullptr, bar(EU_BAR_H2 + c));
          stamp(it, 7 + k);
        }
      }
      mbar_wait(bar(EU_BAR_ACC2B), ph);
      tc_fence_after();
      {
        const int c = 4 + grp;
        uint32_t r[32];
        tmem_ld32(tmem + tlane + EU_COL_ACC2 + 32 * c, r);
        tc_wait_ld();
        convert_chunk(EU_COL_ACC2 + 32 * c, r, nullptr, bar(EU_BAR_H2 + c));
        stamp(it, 9);
      }
      // z buffer it&1: its z was consumed by stage_a0(it) and the previous tile's output was staged in it at the
      // top of this iteration; once the TMA store has read it, refill it with the z of tile it+2
      if (it > 0) bulk_wait_read0();
      request_z(blk + 2 * stride, it & 1u);
      pm_prev = pm_cur;
      blk_prev = blk;
    }
    if (it > 0) epilogue3(blk_prev, it - 1, pm_prev);
    bulk_wait0();   // all output rows are in global memory before the CTA may retire
  } else if (warp == 8 && rank == 0) {
    // ================================ MMA issuer (leader CTA) =====================================
    // The whole warp runs the loop (warp-uniform control flow keeps descriptors in uniform registers);
    // one elected lane issues the MMAs and commits.
    const uint32_t id_w = idesc_f16(256, 192), id_o = idesc_f16(256, 64);
    const uint32_t id_sw = id_w | (1u << 16), id_so = id_o | (1u << 16);   // selector: B operand is MN-major
    const uint32_t id_a = idesc_f16(256, 128);
    const bool d1 = a.drop_terms & 1, d3z = a.drop_terms & 2, d2 = a.drop_terms & 4, d3 = a.drop_terms & 8;
    // Tensor-pipe order per tile n:  A(n) B(n) | 2a(n) 2b(n) 3(n) | A(n+1) B(n+1) ...
    //   A(n): acc1 = S [P;Q] + A0 W1z^T          B(n): acc3 = S [U;V] + A0 Wfz^T  (releases A0 and the U/V tile)
    auto first_layer = [&](uint32_t n) {
      mbar_wait(bar(EU_BAR_PQR), n & 1u);
      mbar_wait(bar(EU_BAR_A0), n & 1u);
      tc_fence_after();
      issue_selector(tmem + EU_COL_ACC1, sbase + EU_SMEM_SEL, sbase + EU_SMEM_PQ, 12, id_sw);
      issue_chunk(tmem + EU_COL_ACC1, tmem + EU_COL_A0, sbase + EU_B1A_HI, sbase + EU_B1A_LO, 0, EU_NL1A, id_w, false, d1);
      issue_chunk(tmem + EU_COL_ACC1, tmem + EU_COL_A0 + 32, sbase + EU_B1A_HI, sbase + EU_B1A_LO, 2, EU_NL1A, id_w, false, d1);
      if (elect_one()) commit_pair(bar(EU_BAR_ACC1));
      __syncwarp();
      // acc3 of the previous tile must have been read by its LayerNorm epilogue
      mbar_wait(bar(EU_BAR_UVR), n & 1u);
      if (n > 0) mbar_wait(bar(EU_BAR_ACC3F), (n - 1) & 1u);
      tc_fence_after();
      issue_selector(tmem + EU_COL_ACC3, sbase + EU_SMEM_SEL, sbase + EU_SMEM_UV, 4, id_so);
      issue_chunk(tmem + EU_COL_ACC3, tmem + EU_COL_A0, sbase + EU_B1B_HI, sbase + EU_B1B_LO, 0, EU_NL1B, id_o, false, d3z);
      issue_chunk(tmem + EU_COL_ACC3, tmem + EU_COL_A0 + 32, sbase + EU_B1B_HI, sbase + EU_B1B_LO, 2, EU_NL1B, id_o, false, d3z);
      if (elect_one()) commit_pair(bar(EU_BAR_PB));
      __syncwarp();
    };
    uint32_t it = 0;
    if (first < nblk) first_layer(0);
    for (uint32_t blk = first; blk < nblk; blk += stride, ++it) {
      const uint32_t ph = it & 1u;
      stamp(it, 0);
      // layer 2 starts from b2 (selector K step 0 x the constant b2 tiles; no dependence on h1)
      issue_bias(tmem + EU_COL_ACC2, sbase + EU_SMEM_SEL, sbase + EU_SMEM_B2, 8, id_a | (1u << 16));
      issue_bias(tmem + EU_COL_ACC2B, sbase + EU_SMEM_SEL, sbase + EU_SMEM_B2 + 4096, 4, id_so);
      // layer 2, output columns 0..127: K chunk by K chunk as epilogue 1 produces h1
      for (int c = 0; c < 6; ++c) {
        mbar_wait(bar(EU_BAR_H1 + c), ph);
        tc_fence_after();
        stamp(it, 2 + c);
        issue_chunk(tmem + EU_COL_ACC2, tmem + EU_COL_ACC1 + 32 * c, sbase + EU_B2A_HI, sbase + EU_B2A_LO, 2 * c, EU_NL2A,
                    id_a, false, d2);
      }
      if (elect_one()) commit_pair(bar(EU_BAR_ACC2A));
      __syncwarp();
      // layer 2, output columns 128..191 (the epilogue of columns 0..127 runs underneath)
      for (int c = 0; c < 6; ++c)
        issue_chunk(tmem + EU_COL_ACC2B, tmem + EU_COL_ACC1 + 32 * c, sbase + EU_B2B_HI, sbase + EU_B2B_LO, 2 * c, EU_NL2B,
                    id_o, false, d2);
      if (elect_one()) commit_pair(bar(EU_BAR_ACC2B));
      __syncwarp();
      stamp(it, 8);
      // layer 3: acc3 += h2 Wf^T
      for (int c = 0; c < 6; ++c) {
        mbar_wait(bar(EU_BAR_H2 + c), ph);
        tc_fence_after();
        stamp(it, 9 + c);
        issue_chunk(tmem + EU_COL_ACC3, tmem + EU_COL_ACC2 + 32 * c, sbase + EU_B3_HI, sbase + EU_B3_LO, 2 * c, EU_NL3,
                    id_o, false, d3);
      }
      if (elect_one()) commit_pair(bar(EU_BAR_ACC3));
      __syncwarp();
      stamp(it, 15);
      // next tile's first layer (its epilogue-3-of-this-tile runs underneath)
      if (blk + stride < nblk) first_layer(it + 1);
      stamp(it, 1);
    }
  } else if (warp == 9 && lane == 0) {
    // ================================ producer of the P/Q and U/V tiles (both CTAs) ================
    const PackLayout lay = a.lay;
    auto load_pq = [&](uint32_t blk) {
      const int b = blk / (T1 * T1), rem = blk % (T1 * T1);
      const size_t slab_b = (size_t)(b * 2 + (int)rank) * lay.KG;
      const size_t oi = (slab_b + 2 * (rem / T1)) * 1536, oj = (slab_b + 2 * (rem % T1)) * 1536;
      const unsigned char* Ph = a.terms + lay.region(0);
      const unsigned char* Qh = a.terms + lay.region(1);
      mbar_arrive_expect_tx(bar(EU_BAR_PQF), 4 * 3072);
      bulk_g2s(sbase + EU_SMEM_PQ, Ph + oi, 3072, bar(EU_BAR_PQF));
      bulk_g2s(sbase + EU_SMEM_PQ + 3072, Qh + oj, 3072, bar(EU_BAR_PQF));
      bulk_g2s(sbase + EU_SMEM_PQ + 6144, Ph + lay.wide_bytes + oi, 3072, bar(EU_BAR_PQF));
      bulk_g2s(sbase + EU_SMEM_PQ + 9216, Qh + lay.wide_bytes + oj, 3072, bar(EU_BAR_PQF));
    };
    auto load_uv = [&](uint32_t blk) {
      const int b = blk / (T1 * T1), rem = blk % (T1 * T1);
      const size_t slab_b = (size_t)(b * 2 + (int)rank) * lay.KG;
      const size_t oi = (slab_b + 2 * (rem / T1)) * 512, oj = (slab_b + 2 * (rem % T1)) * 512;
      const unsigned char* Uh = a.terms + lay.region(2);
      const unsigned char* Vh = a.terms + lay.region(3);
      mbar_arrive_expect_tx(bar(EU_BAR_UVF), 4 * 1024);
      bulk_g2s(sbase + EU_SMEM_UV, Uh + oi, 1024, bar(EU_BAR_UVF));
      bulk_g2s(sbase + EU_SMEM_UV + 1024, Vh + oj, 1024, bar(EU_BAR_UVF));
      bulk_g2s(sbase + EU_SMEM_UV + 2048, Uh + lay.out_bytes + oi, 1024, bar(EU_BAR_UVF));
      bulk_g2s(sbase + EU_SMEM_UV + 3072, Vh + lay.out_bytes + oj, 1024, bar(EU_BAR_UVF));
    };
    if (first < nblk) { load_pq(first); load_uv(first); }
    uint32_t it = 0;
    for (uint32_t blk = first; blk < nblk; blk += stride, ++it) {
      const uint32_t ph = it & 1u;
      mbar_wait(bar(EU_BAR_PQF), ph);
      mbar_arrive_cluster(bar(EU_BAR_PQR), 0);
      mbar_wait(bar(EU_BAR_UVF), ph);
      mbar_arrive_cluster(bar(EU_BAR_UVR), 0);
      const bool more = blk + stride < nblk;
      mbar_wait(bar(EU_BAR_ACC1), ph);          // the P/Q tile has been consumed
      if (more) load_pq(blk + stride);
      mbar_wait(bar(EU_BAR_PB), ph);            // the U/V tile has been consumed
      if (more) load_uv(blk + stride);
    }
  }

  // ---- teardown: nobody leaves (or frees tensor memory) while the pair may still touch this CTA
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 8) tmem_dealloc_pair(tmem, 512);
}

void edge_umma_init() {
  cudaFuncSetAttribute(edge_transition_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EU_SMEM_TOTAL);
  cudaFuncSetAttribute(edge_transition_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EU_SMEM_TOTAL);
}

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// z [B, L, L, 64] fp32 as a 3-D tensor (c: 64, j: L, bi: B*L) with boxes of [32 c, 8 j, 1 bi] = 1 KB, 128-byte
// swizzle.  The encoder lives in libcuda; it is looked up once through the runtime (no link-time dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}
void edge_umma_resolve_driver() { (void)encode_tiled_fn(); }   // at pf_init, not inside a stream capture
int encode_z_map(void* tensor_map, const float* z, int B, int L) {
  CUtensorMap* m = static_cast<CUtensorMap*>(tensor_map);
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return static_cast<int>(cudaErrorNotSupported);
  const cuuint64_t dims[3] = {64, (cuuint64_t)L, (cuuint64_t)B * L};
  const cuuint64_t strides[2] = {256, (cuuint64_t)L * 256};
  const cuuint32_t box[3] = {32, 8, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(z), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PF_OK : static_cast<int>(cudaErrorInvalidValue);
}

size_t edge_umma_pack_bytes(int B, int L) {
  const PackLayout lay = pack_layout(B, L);
  return al256(2 * (size_t)EU_RANK_BYTES) + al256(4 * lay.wide_bytes + 4 * lay.out_bytes);
}

size_t edge_umma_weight_image_bytes() { return 2 * (size_t)EU_RANK_BYTES; }

int launch_edge_umma_pack_weights(const float* w1, const float* w2, const float* wf, void* image, cudaStream_t st) {
  const int n = 2 * EU_RANK_BYTES / 16;
  edge_umma_pack_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(w1, w2, wf, static_cast<uint4*>(image));
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_edge_umma(const float* z_in, const float* P, const float* Q, const float* U, const float* V,
                     const float* w1, const float* w2, const float* wf, const float* b2, const float* ln_g,
                     const float* ln_b, const float* mask, float* z_out, void* wpack, int B, int L, cudaStream_t st,
                     const void* prepacked_weights) {
  unsigned char* terms = static_cast<unsigned char*>(wpack) + al256(2 * (size_t)EU_RANK_BYTES);
  const PackLayout lay = pack_layout(B, L);
  {
    if (!prepacked_weights) PF_TRY(launch_edge_umma_pack_weights(w1, w2, wf, wpack, st));
    const size_t m = 2 * (size_t)B * 2 * lay.KG * 16 * 8;
    edge_umma_pack_terms_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(P, Q, U, V, terms, lay, B, L);
    PF_CHECK_LAUNCH();
  }
  EdgeUArgs a;
  PF_TRY(encode_z_map(&a.tm_in, z_in, B, L));
  PF_TRY(encode_z_map(&a.tm_out, z_out, B, L));
  a.z_in = z_in; a.terms = terms; a.lay = lay; a.b2 = b2; a.ln_g = ln_g; a.ln_b = ln_b; a.mask = mask;
  a.wpack = static_cast<const uint4*>(prepacked_weights ? prepacked_weights : wpack); a.z_out = z_out; a.B = B; a.L = L;
  a.tiles_1d = (L + 15) / 16;
  a.total_blocks = B * a.tiles_1d * a.tiles_1d;
  a.drop_terms = opt_edge_terms();
  int clusters = num_sms() / 2;
  if (clusters > a.total_blocks) clusters = a.total_blocks;
  if (clusters < 1) clusters = 1;
  a.dbg = static_cast<long long*>(debug_buffer(3 * 4 * 16 * sizeof(long long)));
  profile_begin(1, st);
  if (a.dbg) edge_transition_umma_kernel<true><<<2 * clusters, EU_THREADS, EU_SMEM_TOTAL, st>>>(a);
  else edge_transition_umma_kernel<false><<<2 * clusters, EU_THREADS, EU_SMEM_TOTAL, st>>>(a);
  profile_end(1, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

}  // namespace pf
