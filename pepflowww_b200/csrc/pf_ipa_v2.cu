// K3, variant 2 ("ipa_impl" = 2): tensor-core fused invariant-point attention with every operand of a key
// tile resident in shared memory before it is needed (models_con/ipa_pytorch.py:393-473).
//
// Differences to variant 1 (pf_ipa_tc.cu), all aimed at the exposed L2 latency that dominated it:
//   * The point-distance term rides on the tensor core.  -1/2 c_h sum_p |q_p - k_p|^2 =
//     c_h q.k - 1/2 c_h |k|^2 - 1/2 c_h |q|^2: the last term is constant along a softmax row and drops out,
//     the first extends the scalar Q K^T contraction from 128 to 152 channels (Q' = [q/sqrt(3C) | c_h q_pts],
//     K' = [k | k_pts]), the middle one is a per-(key, head) bias computed once by the pack kernel in fp32.
//     (3xFP16 split products; the cancellation is bounded by 2^-22 |q||k| c_h ~ 3e-5 on a logit.)
//   * Keys are streamed in tiles of 8.  The pre-packed K' and V' fragments of all heads for one tile, the key
//     biases and the key mask form ONE contiguous 84 KB blob per (complex, key tile) that a single bulk copy
//     (TMA engine, mbarrier completion) lands in shared memory while the previous tile is finished; the MMA
//     operands are then conflict-free 16 / 8 byte shared-memory loads instead of L2 round trips.
//   * z tiles [16 i x 8 j x 64] keep the 2-stage cp.async ring.
// CTA = 16 query rows x 8 heads; warp roles alternate between pair-major (pair bias W_b z on the tensor core,
// o_pair accumulation) and head-major (Q'K'^T, online softmax, P V with m16n8k8) as in variant 1.
#include <cuda.h>   // CUtensorMap

#include "pf_common.cuh"
#include "pf_split.cuh"
#include "pf_umma.cuh"

namespace pf {

using namespace umma;

constexpr int V2_TQ = 16, V2_TK = 8;
constexpr int V2_ZP = 68;            // padded pair row (floats)
constexpr int V2_KS = 10;            // K steps of Q'K'^T: 128 scalar + 24 point + 8 zero channels
constexpr int V2_VNT = 21;           // n-tiles of [v(128) | v_pts(36) | pad(4)]
constexpr int V2_BLOB_K = H * V2_KS * 32 * 16;    // 40960: [h][ks][lane] uint4 {hi b0, hi b1, lo b0, lo b1}
constexpr int V2_BLOB_V = H * V2_VNT * 32 * 8;    // 43008: [h][nt][lane] uint2 {hi, lo}
constexpr int V2_BLOB_KB = H * V2_TK * 4;         // 256:   [h][key] fp32  -1/2 c_h |k_pts|^2
constexpr int V2_BLOB_M = V2_TK * 4;              // 32:    [key] fp32 residue mask
constexpr int V2_BLOB = V2_BLOB_K + V2_BLOB_V + V2_BLOB_KB + V2_BLOB_M;   // 84256 (multiple of 16)
// blob layout: K' | key bias | key mask | V'   (the first three are needed first and form one bulk copy in v3)
constexpr int V2_OFF_KB = V2_BLOB_K, V2_OFF_M = V2_OFF_KB + V2_BLOB_KB, V2_OFF_V = V2_OFF_M + V2_BLOB_M;
constexpr int V2_BLOB_HEAD = V2_OFF_V;            // 41248 bytes: K' + key bias + mask
static_assert(V2_OFF_V % 16 == 0, "V' part must stay 16-byte aligned");
constexpr int V2_QTILE_U4 = V2_KS * 32 * 2;       // per (b, h, it): [ks][lane]{hi, lo} uint4
constexpr float V2_QSCALE = 0.05103103630798288f; // sqrt(1/(3*128))

struct IpaPack2Args {
  const float* proj; const float* pts; const float* head_w; const float* mask;
  unsigned char* blobs; uint4* Qp;
  int B, L, JT, IT;
};

__device__ __forceinline__ float v2_kprime(const IpaPack2Args& a, size_t row, int h, int kk) {
  if (kk < C) return a.proj[row * NPROJ + OFF_KV + h * 2 * C + kk];
  if (kk < C + PQ * 3) return a.pts[(row * H + h) * (NPT * 3) + PQ * 3 + (kk - C)];
  return 0.f;
}
__device__ __forceinline__ float v2_qprime(const IpaPack2Args& a, size_t row, int h, int kk, float ch) {
  if (kk < C) return a.proj[row * NPROJ + OFF_Q + h * C + kk] * V2_QSCALE;
  if (kk < C + PQ * 3) return a.pts[(row * H + h) * (NPT * 3) + (kk - C)] * ch;
  return 0.f;
}
__device__ __forceinline__ float v2_vprime(const IpaPack2Args& a, size_t row, int h, int n) {
  if (n < C) return a.proj[row * NPROJ + OFF_KV + h * 2 * C + C + n];
  if (n < C + PV * 3) return a.pts[(row * H + h) * (NPT * 3) + 2 * PQ * 3 + (n - C)];
  return 0.f;
}

__global__ void ipa_pack2_kernel(IpaPack2Args a) {
  const int L = a.L;
  const size_t nK = (size_t)a.B * a.JT * H * V2_KS * 32;
  const size_t nV = (size_t)a.B * a.JT * H * V2_VNT * 32;
  const size_t nB = (size_t)a.B * a.JT * (H + 1) * V2_TK;
  const size_t nQ = (size_t)a.B * H * a.IT * V2_KS * 32;
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx < nK) {
    const int lane = idx & 31;
    size_t r = idx >> 5;
    const int ks = r % V2_KS; r /= V2_KS;
    const int h = r % H; r /= H;
    const int jt = r % a.JT, b = (int)(r / a.JT);
    const int g = lane >> 2, t = lane & 3, j = jt * V2_TK + g;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < L) {
      const size_t row = (size_t)b * L + j;
      const int kk = ks * 16 + 2 * t;
      v[0] = v2_kprime(a, row, h, kk); v[1] = v2_kprime(a, row, h, kk + 1);
      v[2] = v2_kprime(a, row, h, kk + 8); v[3] = v2_kprime(a, row, h, kk + 9);
    }
    uint4 o;
    split_pair(v[0], v[1], o.x, o.z);
    split_pair(v[2], v[3], o.y, o.w);
    unsigned char* blob = a.blobs + ((size_t)b * a.JT + jt) * V2_BLOB;
    reinterpret_cast<uint4*>(blob)[(h * V2_KS + ks) * 32 + lane] = o;
    return;
  }
  idx -= nK;
  if (idx < nV) {
    const int lane = idx & 31;
    size_t r = idx >> 5;
    const int nt = r % V2_VNT; r /= V2_VNT;
    const int h = r % H; r /= H;
    const int jt = r % a.JT, b = (int)(r / a.JT);
    const int g = lane >> 2, t = lane & 3, n = nt * 8 + g, j = jt * V2_TK + 2 * t;
    const float v0 = (j < L) ? v2_vprime(a, (size_t)b * L + j, h, n) : 0.f;
    const float v1 = (j + 1 < L) ? v2_vprime(a, (size_t)b * L + j + 1, h, n) : 0.f;
    uint2 o;
    split_pair(v0, v1, o.x, o.y);
    unsigned char* blob = a.blobs + ((size_t)b * a.JT + jt) * V2_BLOB + V2_OFF_V;
    reinterpret_cast<uint2*>(blob)[(h * V2_VNT + nt) * 32 + lane] = o;
    return;
  }
  idx -= nV;
  if (idx < nB) {
    const int key = idx % V2_TK;
    size_t r = idx / V2_TK;
    const int h = r % (H + 1); r /= (H + 1);
    const int jt = r % a.JT, b = (int)(r / a.JT);
    const int j = jt * V2_TK + key;
    float* dst = reinterpret_cast<float*>(a.blobs + ((size_t)b * a.JT + jt) * V2_BLOB + V2_OFF_KB);
    if (h == H) {
      dst[H * V2_TK + key] = (j < L) ? a.mask[(size_t)b * L + j] : 0.f;
    } else {
      float s = 0.f;
      if (j < L) {
        const float* kp = a.pts + (((size_t)b * L + j) * H + h) * (NPT * 3) + PQ * 3;
#pragma unroll
        for (int e = 0; e < PQ * 3; ++e) s = fmaf(kp[e], kp[e], s);
      }
      dst[h * V2_TK + key] = -0.5f * a.head_w[h] * s;
    }
    return;
  }
  idx -= nB;
  if (idx < nQ) {
    const int lane = idx & 31;
    size_t r = idx >> 5;
    const int ks = r % V2_KS; r /= V2_KS;
    const int it = r % a.IT; r /= a.IT;
    const int h = r % H, b = (int)(r / H);
    const int g = lane >> 2, t = lane & 3;
    const float ch = a.head_w[h];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = it * V2_TQ + g + half * 8;
      if (i < L) {
        const size_t row = (size_t)b * L + i;
        const int kk = ks * 16 + 2 * t;
        v[half * 2 + 0] = v2_qprime(a, row, h, kk, ch); v[half * 2 + 1] = v2_qprime(a, row, h, kk + 1, ch);
        v[4 + half * 2 + 0] = v2_qprime(a, row, h, kk + 8, ch); v[4 + half * 2 + 1] = v2_qprime(a, row, h, kk + 9, ch);
      }
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x);
    split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z);
    split_pair(v[6], v[7], hi.w, lo.w);
    uint4* q = a.Qp + ((((size_t)b * H + h) * a.IT + it) * V2_KS + ks) * 64 + lane * 2;
    q[0] = hi;
    q[1] = lo;
  }
}

// D(16x8, fp32) += A(16x8, fp16, row) * B(8x8, fp16, col)
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void v2_cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}

// ---- shared memory (bytes)
constexpr int V2_SM_ZSTAGE = V2_TQ * V2_TK * V2_ZP * 4;      // 34816
constexpr int V2_SM_Z = 0;
constexpr int V2_SM_BLOB = V2_SM_Z + 2 * V2_SM_ZSTAGE;       // 69632
constexpr int V2_SM_OP = V2_SM_BLOB + V2_BLOB;               // o_pair_raw [16 i][8 h][64] fp32
constexpr int V2_BP = 10;                                    // pitch of the bias rows [h][i][10]
constexpr int V2_SM_BIAS = V2_SM_OP + V2_TQ * H * CZ * 4;
constexpr int V2_PP = 68;                                    // pitch of the P rows [i][key 8][h 8] (+4)
constexpr int V2_SM_P = V2_SM_BIAS + H * V2_TQ * V2_BP * 4;
constexpr int V2_SM_ALPHA = V2_SM_P + V2_TQ * V2_PP * 4;     // [h][i]
constexpr int V2_SM_L = V2_SM_ALPHA + H * V2_TQ * 4;         // [h][i] final 1/row sums
constexpr int V2_SM_BAR = V2_SM_L + H * V2_TQ * 4;
constexpr int V2_SMEM = V2_SM_BAR + 16;
static_assert(V2_SM_BLOB % 16 == 0 && V2_SM_OP % 16 == 0 && V2_SM_P % 16 == 0, "alignment");
static_assert(V2_SMEM <= 232448, "shared memory budget");

struct Ipa2Args {
  IpaArgs a;
  const unsigned char* blobs; const uint4* Qp;
  int JT, IT;
};

__global__ void __launch_bounds__(256, 1) ipa_attention_v2_kernel(Ipa2Args p) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* zs = reinterpret_cast<float*>(smem + V2_SM_Z);
  const unsigned char* blob = smem + V2_SM_BLOB;
  float* sop = reinterpret_cast<float*>(smem + V2_SM_OP);
  float* sbias = reinterpret_cast<float*>(smem + V2_SM_BIAS);
  float* sP = reinterpret_cast<float*>(smem + V2_SM_P);
  float* salpha = reinterpret_cast<float*>(smem + V2_SM_ALPHA);
  float* sl = reinterpret_cast<float*>(smem + V2_SM_L);
  const uint32_t bar = smem_u32(smem + V2_SM_BAR);
  const IpaArgs& a = p.a;
  const int L = a.L;
  const int b = blockIdx.y, it = blockIdx.x, i0 = it * V2_TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const size_t rowb = (size_t)b * L;
  const int h = warp;                                  // head-major role
  const float sc_b = 0.5773502691896257f;              // sqrt(1/3)

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  for (int idx = tid; idx < V2_TQ * H * CZ; idx += 256) sop[idx] = 0.f;
  // W_b as B fragments (pair-bias mma): rows n = head (8), k = channel; sqrt(1/3) folded in
  uint32_t wbh[4][2], wbl[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float* w = a.w_b + g * CZ + ks * 16 + 2 * t;
    split_pair(w[0] * sc_b, w[1] * sc_b, wbh[ks][0], wbl[ks][0]);
    split_pair(w[8] * sc_b, w[9] * sc_b, wbh[ks][1], wbl[ks][1]);
  }
  // Q' fragments of head h (hi / lo)
  uint4 qh[V2_KS], ql[V2_KS];
  {
    const uint4* qp = p.Qp + (((size_t)b * H + h) * p.IT + it) * V2_QTILE_U4;
#pragma unroll
    for (int ks = 0; ks < V2_KS; ++ks) {
      qh[ks] = qp[(ks * 32 + lane) * 2];
      ql[ks] = qp[(ks * 32 + lane) * 2 + 1];
    }
  }
  const float bbias = sc_b * a.b_b[h];
  const float mi_lo = (i0 + g < L) ? a.mask[rowb + i0 + g] : 0.f;
  const float mi_hi = (i0 + g + 8 < L) ? a.mask[rowb + i0 + g + 8] : 0.f;

  float O[V2_VNT][4];
#pragma unroll
  for (int n = 0; n < V2_VNT; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  auto issue_z = [&](int jt) {
    const int j0 = jt * V2_TK;
    float* zd = zs + (jt & 1) * (V2_SM_ZSTAGE / 4);
    for (int c = tid; c < V2_TQ * V2_TK * 16; c += 256) {          // 16-byte chunks: pair (i,j) x 16 chunks
      const int ch = c & 15, pr = c >> 4, j = pr & 7, i = pr >> 3;
      const bool ok = (i0 + i < L) && (j0 + j < L);
      const float* src = a.z + (((rowb + (ok ? i0 + i : 0)) * L + (ok ? j0 + j : 0)) * CZ) + ch * 4;
      v2_cp_async16_zfill(zd + pr * V2_ZP + ch * 4, src, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  auto issue_blob = [&](int jt) {   // one thread; the whole tile's K' / V' / key bias / mask in one bulk copy
    mbar_arrive_expect_tx(bar, V2_BLOB);
    bulk_g2s(smem_u32(smem + V2_SM_BLOB), p.blobs + ((size_t)b * p.JT + jt) * V2_BLOB, V2_BLOB, bar);
  };

  __syncthreads();                                     // barrier initialised, sop zeroed
  issue_z(0);
  if (tid == 0) issue_blob(0);

  const float* skb = reinterpret_cast<const float*>(blob + V2_OFF_KB);   // [h][8] then mask [8]
  for (int jt = 0; jt < p.JT; ++jt) {
    const int j0 = jt * V2_TK;
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();                                   // z tile jt visible; everyone finished tile jt-1
    if (jt + 1 < p.JT) issue_z(jt + 1);
    const float* zt = zs + (jt & 1) * (V2_SM_ZSTAGE / 4);

    // ================= phase 1 (pair-major): pair bias for rows 2w, 2w+1 x 8 keys, all heads
    {
      float acc[2][4];                                 // small terms | hi*hi: two independent chains
      acc[0][0] = acc[0][1] = acc[0][2] = acc[0][3] = 0.f;
      acc[1][0] = acc[1][1] = acc[1][2] = acc[1][3] = 0.f;
      const float* zr_lo = zt + ((2 * warp) * V2_TK + g) * V2_ZP;       // pair (row 2w,   key g)
      const float* zr_hi = zt + ((2 * warp + 1) * V2_TK + g) * V2_ZP;   // pair (row 2w+1, key g)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int c = ks * 16 + 2 * t;
        const float2 x0 = *reinterpret_cast<const float2*>(zr_lo + c);
        const float2 x1 = *reinterpret_cast<const float2*>(zr_hi + c);
        const float2 x2 = *reinterpret_cast<const float2*>(zr_lo + c + 8);
        const float2 x3 = *reinterpret_cast<const float2*>(zr_hi + c + 8);
        uint32_t ah[4], al[4];
        split_pair(x0.x, x0.y, ah[0], al[0]);
        split_pair(x1.x, x1.y, ah[1], al[1]);
        split_pair(x2.x, x2.y, ah[2], al[2]);
        split_pair(x3.x, x3.y, ah[3], al[3]);
        mma16816(acc[0], al, wbh[ks][0], wbh[ks][1]);
        mma16816(acc[1], ah, wbh[ks][0], wbh[ks][1]);
        mma16816(acc[0], ah, wbl[ks][0], wbl[ks][1]);
      }
      // C: (pair g -> row 2w, key g; heads 2t, 2t+1), (pair g+8 -> row 2w+1, key g)
      sbias[((2 * t) * V2_TQ + 2 * warp) * V2_BP + g] = acc[0][0] + acc[1][0];
      sbias[((2 * t + 1) * V2_TQ + 2 * warp) * V2_BP + g] = acc[0][1] + acc[1][1];
      sbias[((2 * t) * V2_TQ + 2 * warp + 1) * V2_BP + g] = acc[0][2] + acc[1][2];
      sbias[((2 * t + 1) * V2_TQ + 2 * warp + 1) * V2_BP + g] = acc[0][3] + acc[1][3];
    }
    __syncthreads();                                   // (A) bias tile complete
    mbar_wait(bar, jt & 1);                            // K' / V' / key bias / mask of this tile have landed

    // ================= phase 2 (head-major): S = Q' K'^T for head h, 16 rows x 8 keys
    float S[4];
    {
      float Sa[4] = {0.f, 0.f, 0.f, 0.f}, Sb[4] = {0.f, 0.f, 0.f, 0.f};
      const uint4* kp = reinterpret_cast<const uint4*>(blob) + (h * V2_KS) * 32 + lane;
#pragma unroll
      for (int ks = 0; ks < V2_KS; ++ks) {
        const uint4 k0 = kp[ks * 32];
        const uint32_t ah[4] = {qh[ks].x, qh[ks].y, qh[ks].z, qh[ks].w};
        const uint32_t al[4] = {ql[ks].x, ql[ks].y, ql[ks].z, ql[ks].w};
        mma16816(Sa, al, k0.x, k0.y);
        mma16816(Sb, ah, k0.x, k0.y);
        mma16816(Sa, ah, k0.z, k0.w);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) S[e] = Sa[e] + Sb[e];
    }
    // ================= phase 3: logits, online softmax (row g: S[0..1], row g+8: S[2..3]; keys 2t, 2t+1)
    {
      float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 2 * t + e;
        const bool valid = (j0 + j < L);
        const float mj = skb[H * V2_TK + j];
        const float kb = skb[h * V2_TK + j] + bbias;
        const float b_lo = sbias[(h * V2_TQ + g) * V2_BP + j], b_hi = sbias[(h * V2_TQ + g + 8) * V2_BP + j];
        float x_lo = S[e] + b_lo + kb + 1e5f * (mi_lo * mj - 1.f);
        float x_hi = S[2 + e] + b_hi + kb + 1e5f * (mi_hi * mj - 1.f);
        x_lo = valid ? x_lo : -INFINITY;
        x_hi = valid ? x_hi : -INFINITY;
        S[e] = x_lo; S[2 + e] = x_hi;
        mx_lo = fmaxf(mx_lo, x_lo); mx_hi = fmaxf(mx_hi, x_hi);
      }
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
      mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
      mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
      const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
      const float al_lo = expf(m_lo - mn_lo), al_hi = expf(m_hi - mn_hi);   // exp(-inf) = 0 on the first tile
      m_lo = mn_lo; m_hi = mn_hi;
      float ps_lo = 0.f, ps_hi = 0.f;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float p_lo = expf(S[e] - mn_lo), p_hi = expf(S[2 + e] - mn_hi);
        S[e] = p_lo; S[2 + e] = p_hi;
        ps_lo += p_lo; ps_hi += p_hi;
        const int j = 2 * t + e;
        sP[g * V2_PP + j * 8 + h] = p_lo;
        sP[(g + 8) * V2_PP + j * 8 + h] = p_hi;
      }
      l_lo = l_lo * al_lo + ps_lo;
      l_hi = l_hi * al_hi + ps_hi;
      if (t == 0) { salpha[h * V2_TQ + g] = al_lo; salpha[h * V2_TQ + g + 8] = al_hi; }
#pragma unroll
      for (int n = 0; n < V2_VNT; ++n) { O[n][0] *= al_lo; O[n][1] *= al_lo; O[n][2] *= al_hi; O[n][3] *= al_hi; }
    }
    // ================= phase 4: O += P [V | v_pts]   (m16n8k8: K = the 8 keys of the tile)
    {
      uint32_t ph0, pl0, ph1, pl1;
      split_pair(S[0], S[1], ph0, pl0);                // a0: row g,   keys 2t, 2t+1
      split_pair(S[2], S[3], ph1, pl1);                // a1: row g+8
      const uint2* vp = reinterpret_cast<const uint2*>(blob + V2_OFF_V) + (h * V2_VNT) * 32 + lane;
#pragma unroll
      for (int n = 0; n < V2_VNT; ++n) {
        const uint2 v = vp[n * 32];
        mma1688(O[n], pl0, pl1, v.x);
        mma1688(O[n], ph0, ph1, v.y);
        mma1688(O[n], ph0, ph1, v.x);
      }
    }
    __syncthreads();                                   // (B) P tile and alpha complete; the blob has been consumed
    if (tid == 0 && jt + 1 < p.JT) issue_blob(jt + 1);

    // ================= phase 5 (pair-major): o_pair_raw[i, h, :] = alpha * old + sum_j P z
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int i = 2 * warp + mt;
      float acc[H][2];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        const float al = salpha[hh * V2_TQ + i];
        const float2 o = *reinterpret_cast<const float2*>(sop + (i * H + hh) * CZ + 2 * lane);
        acc[hh][0] = o.x * al; acc[hh][1] = o.y * al;
      }
#pragma unroll
      for (int j = 0; j < V2_TK; ++j) {
        const float2 zv = *reinterpret_cast<const float2*>(zt + (i * V2_TK + j) * V2_ZP + 2 * lane);
        const float4 p0 = *reinterpret_cast<const float4*>(sP + i * V2_PP + j * 8);
        const float4 p1 = *reinterpret_cast<const float4*>(sP + i * V2_PP + j * 8 + 4);
        const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
          acc[hh][0] = fmaf(pv[hh], zv.x, acc[hh][0]);
          acc[hh][1] = fmaf(pv[hh], zv.y, acc[hh][1]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < H; ++hh)
        *reinterpret_cast<float2*>(sop + (i * H + hh) * CZ + 2 * lane) = make_float2(acc[hh][0], acc[hh][1]);
    }
  }

  // ================= epilogue
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float il_lo = 1.0f / l_lo, il_hi = 1.0f / l_hi;
  if (t == 0) { sl[h * V2_TQ + g] = il_lo; sl[h * V2_TQ + g + 8] = il_hi; }
  __syncthreads();                                     // last tile's phase 5 done everywhere; zs is free
  float* spt = zs;                                      // [8 h][16 i][40] normalised global-frame o_pt
  {
    const int i_lo = i0 + g, i_hi = i0 + g + 8;
#pragma unroll
    for (int n = 0; n < 16; ++n) {                      // o: channels 8n + 2t, +1
      if (i_lo < L)
        *reinterpret_cast<float2*>(a.feats + (rowb + i_lo) * NFEAT + h * C + n * 8 + 2 * t) =
            make_float2(O[n][0] * il_lo, O[n][1] * il_lo);
      if (i_hi < L)
        *reinterpret_cast<float2*>(a.feats + (rowb + i_hi) * NFEAT + h * C + n * 8 + 2 * t) =
            make_float2(O[n][2] * il_hi, O[n][3] * il_hi);
    }
#pragma unroll
    for (int n = 16; n < V2_VNT; ++n) {
      const int c = (n - 16) * 8 + 2 * t;
      spt[(h * V2_TQ + g) * 40 + c] = O[n][0] * il_lo;
      spt[(h * V2_TQ + g) * 40 + c + 1] = O[n][1] * il_lo;
      spt[(h * V2_TQ + g + 8) * 40 + c] = O[n][2] * il_hi;
      spt[(h * V2_TQ + g + 8) * 40 + c + 1] = O[n][3] * il_hi;
    }
  }
  __syncthreads();
  // o_pt: global -> local frame, norms (ipa_pytorch.py:455-460)
  for (int idx = tid; idx < H * V2_TQ * PV; idx += 256) {
    const int pnt = idx % PV, i = (idx / PV) % V2_TQ, hh = idx / (PV * V2_TQ);
    if (i0 + i >= L) continue;
    const float* R = a.rot + (rowb + i0 + i) * 9;
    const float* tr = a.trans + (rowb + i0 + i) * 3;
    const float* s = spt + (hh * V2_TQ + i) * 40 + pnt * 3;
    const float gx = s[0] - tr[0], gy = s[1] - tr[1], gz = s[2] - tr[2];
    const float lx = R[0] * gx + R[3] * gy + R[6] * gz;
    const float ly = R[1] * gx + R[4] * gy + R[7] * gz;
    const float lz = R[2] * gx + R[5] * gy + R[8] * gz;
    float* f = a.feats + (rowb + i0 + i) * NFEAT + 1024;
    f[0 * 96 + hh * PV + pnt] = lx;
    f[1 * 96 + hh * PV + pnt] = ly;
    f[2 * 96 + hh * PV + pnt] = lz;
    f[3 * 96 + hh * PV + pnt] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
  }
  // o_pair: down_z on the normalised a-weighted pair row (ipa_pytorch.py:469-473); W_dz staged in shared memory
  __syncthreads();
  float* swz = zs;                                      // [16 d][68]
  for (int idx = tid; idx < 16 * CZ; idx += 256) swz[(idx >> 6) * V2_ZP + (idx & 63)] = a.w_dz[idx];
  __syncthreads();
  {
    const int pairidx = tid >> 1, d0 = (tid & 1) * 8;   // pairidx = i * 8 + head
    const int i = pairidx >> 3, hh = pairidx & 7;
    if (i0 + i < L) {
      const float* src = sop + (i * H + hh) * CZ;
      float acc[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[d] = 0.f;
#pragma unroll 4
      for (int c = 0; c < CZ; c += 4) {
        const float4 x = *reinterpret_cast<const float4*>(src + c);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const float4 w = *reinterpret_cast<const float4*>(swz + (d0 + d) * V2_ZP + c);
          acc[d] = fmaf(w.x, x.x, acc[d]);
          acc[d] = fmaf(w.y, x.y, acc[d]);
          acc[d] = fmaf(w.z, x.z, acc[d]);
          acc[d] = fmaf(w.w, x.w, acc[d]);
        }
      }
      const float inv = sl[hh * V2_TQ + i];
      float* out = a.feats + (rowb + i0 + i) * NFEAT + 1024 + 384 + hh * 16 + d0;
      float4 o0, o1;
      o0.x = acc[0] * inv + a.b_dz[d0 + 0]; o0.y = acc[1] * inv + a.b_dz[d0 + 1];
      o0.z = acc[2] * inv + a.b_dz[d0 + 2]; o0.w = acc[3] * inv + a.b_dz[d0 + 3];
      o1.x = acc[4] * inv + a.b_dz[d0 + 4]; o1.y = acc[5] * inv + a.b_dz[d0 + 5];
      o1.z = acc[6] * inv + a.b_dz[d0 + 6]; o1.w = acc[7] * inv + a.b_dz[d0 + 7];
      *reinterpret_cast<float4*>(out) = o0;
      *reinterpret_cast<float4*>(out + 4) = o1;
    }
  }
}

size_t ipa_v2_workspace_bytes(int B, int L) {
  const size_t JT = (L + V2_TK - 1) / V2_TK, IT = (L + V2_TQ - 1) / V2_TQ;
  return (size_t)B * JT * V2_BLOB + (size_t)B * H * IT * V2_QTILE_U4 * sizeof(uint4) + 1024;
}

// =================================================================================================
// variant 3 ("ipa_impl" = 3): the same arithmetic as variant 2, warp-specialised.
//   warps 0-7  (head warps, warp = head):  Q'K'^T -> logits -> online softmax -> P V        (192 registers)
//   warps 8-15 (pair warps, warp = 2 query rows): private 2-stage z ring fed by tensor-map TMA boxes, pair
//              bias on the tensor core, o_pair accumulation as (z^T P^T) MMAs with the running sums held in
//              registers                                                                      (72 registers)
// The two groups run different phases of a key tile at the same time and meet only at two named barriers per
// tile (bias ready: pair -> head, P ready: head -> pair); bias / P / alpha tiles are double buffered by tile
// parity.  K' (+ key bias, mask) and V' are separate bulk copies with their own full / free mbarriers, so the
// next tile's K' lands while this tile's softmax and P V run, and V' while the next Q'K'^T runs.
// setmaxnreg moves registers from the pair warps to the head warps (16 warps x 128 = 8 x 184 + 8 x 72).
constexpr int V3_THREADS = 512;
constexpr int V3_ZSTAGE_W = 4096;                              // one pair warp, one stage: [2 rows][2 halves] TMA boxes
                                                               // of [8 keys][32 ch] (1 KB, 128-byte swizzle)
constexpr int V3_SM_BLOB = 0;                                  // 84256
constexpr int V3_SM_Z = (V3_SM_BLOB + V2_BLOB + 1023) & ~1023; // [8 warps][2 stages][4096]
constexpr int V3_SM_QLO = V3_SM_Z + 8 * 2 * V3_ZSTAGE_W;       // [8 h][10 ks][32 lanes] uint4
constexpr int V3_SM_WB = V3_SM_QLO + H * V2_KS * 32 * 16;      // [4 ks][32 lanes] uint4 {hi b0, hi b1, lo b0, lo b1}
constexpr int V3_SM_BIAS = V3_SM_WB + 4 * 32 * 16;             // [2][8 h][16 i][10]
constexpr int V3_SM_P = V3_SM_BIAS + 2 * H * V2_TQ * V2_BP * 4;   // [2][16 i][68]
constexpr int V3_SM_ALPHA = V3_SM_P + 2 * V2_TQ * V2_PP * 4;   // [2][8 h][16 i]
constexpr int V3_SM_L = V3_SM_ALPHA + 2 * H * V2_TQ * 4;       // [8 h][16 i]
constexpr int V3_SM_BAR = V3_SM_L + H * V2_TQ * 4;             // 4 blob mbarriers + [8 warps][2 stages] z mbarriers
constexpr int V3_SMEM = V3_SM_BAR + 32 + 128;
static_assert(V3_SM_Z % 1024 == 0 && V3_SM_QLO % 16 == 0 && V3_SM_P % 16 == 0 && V3_SM_BAR % 8 == 0, "alignment");
static_assert(V3_SMEM <= 232448, "shared memory budget");
// epilogue re-use of the loop buffers
static_assert(V2_TQ * H * CZ * 4 <= V2_BLOB, "o_pair_raw fits in the blob region");
static_assert(H * V2_TQ * 40 * 4 <= 8 * 2 * V3_ZSTAGE_W, "o_pt scratch fits in the z region");

__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}

struct alignas(64) Ipa3Args {
  CUtensorMap tm_z;             // z as [B*L, L, 64] fp32, box [1, 8, 32], 128-byte swizzle
  Ipa2Args p;
};

// byte offset of z[key][c] inside one pair warp's stage for its query row r (0/1): swizzled TMA boxes
__device__ __forceinline__ int v3_zoff(int r, int key, int c) {
  return ((r * 2 + (c >> 5)) << 10) + (key << 7) + ((((c & 31) >> 2) ^ key) << 4) + ((c & 3) << 2);
}

__global__ void __launch_bounds__(V3_THREADS, 1) ipa_attention_v3_kernel(const __grid_constant__ Ipa3Args args) {
  const Ipa2Args& p = args.p;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned char* blob = smem + V3_SM_BLOB;
  float* sbias = reinterpret_cast<float*>(smem + V3_SM_BIAS);
  float* sP = reinterpret_cast<float*>(smem + V3_SM_P);
  float* salpha = reinterpret_cast<float*>(smem + V3_SM_ALPHA);
  float* sl = reinterpret_cast<float*>(smem + V3_SM_L);
  const uint32_t bar_kfull = smem_u32(smem + V3_SM_BAR), bar_vfull = bar_kfull + 8, bar_kfree = bar_kfull + 16,
                 bar_vfree = bar_kfull + 24;
  const IpaArgs& a = p.a;
  const int L = a.L, JT = p.JT;
  const int b = blockIdx.y, it = blockIdx.x, i0 = it * V2_TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const size_t rowb = (size_t)b * L;
  const float sc_b = 0.5773502691896257f;              // sqrt(1/3)
  const unsigned char* gblob = p.blobs + (size_t)b * JT * V2_BLOB;

  if (tid == 0) {
    mbar_init(bar_kfull, 1);
    mbar_init(bar_vfull, 1);
    mbar_init(bar_kfree, 8);
    mbar_init(bar_vfree, 8);
    for (int q = 0; q < 16; ++q) mbar_init(bar_kfull + 32 + 8 * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tma_prefetch_desc(&args.tm_z);
  }
  // W_b as B fragments of the pair-bias MMA (rows n = head, k = channel; sqrt(1/3) folded in)
  if (tid < 128) {
    const int ks = tid >> 5, gg = (tid & 31) >> 2, tt = tid & 3;
    const float* w = a.w_b + gg * CZ + ks * 16 + 2 * tt;
    uint4 o;
    split_pair(w[0] * sc_b, w[1] * sc_b, o.x, o.z);
    split_pair(w[8] * sc_b, w[9] * sc_b, o.y, o.w);
    reinterpret_cast<uint4*>(smem + V3_SM_WB)[tid] = o;
  }
  // lo halves of the Q' fragments (the hi halves stay in the head warps' registers)
  for (int idx = tid; idx < H * V2_KS * 32; idx += V3_THREADS) {
    const int hh = idx / (V2_KS * 32), rem = idx % (V2_KS * 32);
    reinterpret_cast<uint4*>(smem + V3_SM_QLO)[idx] =
        p.Qp[(((size_t)b * H + hh) * p.IT + it) * V2_QTILE_U4 + rem * 2 + 1];
  }
  __syncthreads();

  if (warp < 8) {
    // ========================================== head warps ======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;\n");
    const int h = warp;
    uint4 qh[V2_KS];
    {
      const uint4* qp = p.Qp + (((size_t)b * H + h) * p.IT + it) * V2_QTILE_U4;
#pragma unroll
      for (int ks = 0; ks < V2_KS; ++ks) qh[ks] = qp[(ks * 32 + lane) * 2];
    }
    const uint4* qlo = reinterpret_cast<const uint4*>(smem + V3_SM_QLO) + (h * V2_KS) * 32 + lane;
    const float bbias = sc_b * a.b_b[h];
    const float mi_lo = (i0 + g < L) ? a.mask[rowb + i0 + g] : 0.f;
    const float mi_hi = (i0 + g + 8 < L) ? a.mask[rowb + i0 + g + 8] : 0.f;
    const float* skb = reinterpret_cast<const float*>(blob + V2_OFF_KB);   // [h][8] then mask [8]
    float O[V2_VNT][4];
#pragma unroll
    for (int n = 0; n < V2_VNT; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_kfull, V2_BLOB_HEAD);
      bulk_g2s(smem_u32(smem + V3_SM_BLOB), gblob, V2_BLOB_HEAD, bar_kfull);
      mbar_arrive_expect_tx(bar_vfull, V2_BLOB_V);
      bulk_g2s(smem_u32(smem + V3_SM_BLOB + V2_OFF_V), gblob + V2_OFF_V, V2_BLOB_V, bar_vfull);
    }
    for (int jt = 0; jt < JT; ++jt) {
      const int j0 = jt * V2_TK, par = jt & 1;
      // ---- S = Q' K'^T (16 rows x 8 keys)
      mbar_wait(bar_kfull, par);
      float S[4];
      float kbv[2], mjv[2];
      {
        float Sa[4] = {0.f, 0.f, 0.f, 0.f}, Sb[4] = {0.f, 0.f, 0.f, 0.f};
        const uint4* kp = reinterpret_cast<const uint4*>(blob) + (h * V2_KS) * 32 + lane;
#pragma unroll
        for (int ks = 0; ks < V2_KS; ++ks) {
          const uint4 k0 = kp[ks * 32];
          const uint4 q_lo = qlo[ks * 32];
          const uint32_t ah[4] = {qh[ks].x, qh[ks].y, qh[ks].z, qh[ks].w};
          const uint32_t al[4] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w};
          mma16816(Sa, al, k0.x, k0.y);
          mma16816(Sb, ah, k0.x, k0.y);
          mma16816(Sa, ah, k0.z, k0.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) S[e] = Sa[e] + Sb[e];
#pragma unroll
        for (int e = 0; e < 2; ++e) { kbv[e] = skb[h * V2_TK + 2 * t + e] + bbias; mjv[e] = skb[H * V2_TK + 2 * t + e]; }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_kfree);          // this warp is done with K' / key bias / mask of the tile
      if (tid == 0 && jt + 1 < JT) {                   // producer: refill the K' part once every head warp is done
        mbar_wait(bar_kfree, par);
        mbar_arrive_expect_tx(bar_kfull, V2_BLOB_HEAD);
        bulk_g2s(smem_u32(smem + V3_SM_BLOB), gblob + (size_t)(jt + 1) * V2_BLOB, V2_BLOB_HEAD, bar_kfull);
      }
      __syncwarp();
      // ---- logits and online softmax (row g: S[0..1], row g+8: S[2..3]; keys 2t, 2t+1)
      named_sync(1 + par, V3_THREADS);                 // pair bias of this tile is in sbias[par]
      {
        const float* bs = sbias + par * (H * V2_TQ * V2_BP);
        float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 2 * t + e;
          const bool valid = (j0 + j < L);
          const float b_lo = bs[(h * V2_TQ + g) * V2_BP + j], b_hi = bs[(h * V2_TQ + g + 8) * V2_BP + j];
          float x_lo = S[e] + b_lo + kbv[e] + 1e5f * (mi_lo * mjv[e] - 1.f);
          float x_hi = S[2 + e] + b_hi + kbv[e] + 1e5f * (mi_hi * mjv[e] - 1.f);
          x_lo = valid ? x_lo : -INFINITY;
          x_hi = valid ? x_hi : -INFINITY;
          S[e] = x_lo; S[2 + e] = x_hi;
          mx_lo = fmaxf(mx_lo, x_lo); mx_hi = fmaxf(mx_hi, x_hi);
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
        const float al_lo = expf(m_lo - mn_lo), al_hi = expf(m_hi - mn_hi);   // exp(-inf) = 0 on the first tile
        m_lo = mn_lo; m_hi = mn_hi;
        float ps_lo = 0.f, ps_hi = 0.f;
        float* Pw = sP + par * (V2_TQ * V2_PP);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p_lo = expf(S[e] - mn_lo), p_hi = expf(S[2 + e] - mn_hi);
          S[e] = p_lo; S[2 + e] = p_hi;
          ps_lo += p_lo; ps_hi += p_hi;
          const int j = 2 * t + e;
          Pw[g * V2_PP + j * 8 + h] = p_lo;
          Pw[(g + 8) * V2_PP + j * 8 + h] = p_hi;
        }
        l_lo = l_lo * al_lo + ps_lo;
        l_hi = l_hi * al_hi + ps_hi;
        if (t == 0) {
          salpha[par * (H * V2_TQ) + h * V2_TQ + g] = al_lo;
          salpha[par * (H * V2_TQ) + h * V2_TQ + g + 8] = al_hi;
        }
        named_arrive(3 + par, V3_THREADS);             // P / alpha of this tile are in sP[par], salpha[par]
#pragma unroll
        for (int n = 0; n < V2_VNT; ++n) { O[n][0] *= al_lo; O[n][1] *= al_lo; O[n][2] *= al_hi; O[n][3] *= al_hi; }
      }
      // ---- O += P [V | v_pts]   (m16n8k8: K = the 8 keys of the tile)
      mbar_wait(bar_vfull, par);
      {
        uint32_t ph0, pl0, ph1, pl1;
        split_pair(S[0], S[1], ph0, pl0);              // a0: row g,   keys 2t, 2t+1
        split_pair(S[2], S[3], ph1, pl1);              // a1: row g+8
        const uint2* vp = reinterpret_cast<const uint2*>(blob + V2_OFF_V) + (h * V2_VNT) * 32 + lane;
#pragma unroll
        for (int n = 0; n < V2_VNT; ++n) {
          const uint2 v = vp[n * 32];
          mma1688(O[n], pl0, pl1, v.x);
          mma1688(O[n], ph0, ph1, v.y);
          mma1688(O[n], ph0, ph1, v.x);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_vfree);
      if (tid == 0 && jt + 1 < JT) {
        mbar_wait(bar_vfree, par);
        mbar_arrive_expect_tx(bar_vfull, V2_BLOB_V);
        bulk_g2s(smem_u32(smem + V3_SM_BLOB + V2_OFF_V), gblob + (size_t)(jt + 1) * V2_BLOB + V2_OFF_V, V2_BLOB_V,
                 bar_vfull);
      }
      __syncwarp();
    }
    // ---- head epilogue: normalise, write o, stage o_pt
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float il_lo = 1.0f / l_lo, il_hi = 1.0f / l_hi;
    if (t == 0) { sl[h * V2_TQ + g] = il_lo; sl[h * V2_TQ + g + 8] = il_hi; }
    named_sync(5, V3_THREADS);                         // every warp has left the tile loop: z region and blob are free
    float* spt = reinterpret_cast<float*>(smem + V3_SM_Z);   // [8 h][16 i][40] normalised global-frame o_pt
    const int i_lo = i0 + g, i_hi = i0 + g + 8;
#pragma unroll
    for (int n = 0; n < 16; ++n) {                     // o: channels 8n + 2t, +1
      if (i_lo < L)
        *reinterpret_cast<float2*>(a.feats + (rowb + i_lo) * NFEAT + h * C + n * 8 + 2 * t) =
            make_float2(O[n][0] * il_lo, O[n][1] * il_lo);
      if (i_hi < L)
        *reinterpret_cast<float2*>(a.feats + (rowb + i_hi) * NFEAT + h * C + n * 8 + 2 * t) =
            make_float2(O[n][2] * il_hi, O[n][3] * il_hi);
    }
#pragma unroll
    for (int n = 16; n < V2_VNT; ++n) {
      const int c = (n - 16) * 8 + 2 * t;
      spt[(h * V2_TQ + g) * 40 + c] = O[n][0] * il_lo;
      spt[(h * V2_TQ + g) * 40 + c + 1] = O[n][1] * il_lo;
      spt[(h * V2_TQ + g + 8) * 40 + c] = O[n][2] * il_hi;
      spt[(h * V2_TQ + g + 8) * 40 + c + 1] = O[n][3] * il_hi;
    }
    asm volatile("setmaxnreg.dec.sync.aligned.u32 128;\n");
  } else {
    // ========================================== pair warps ======================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;\n");
    const int pw = warp - 8, r0 = 2 * pw;             // this warp owns query rows r0, r0 + 1
    const uint32_t zbase = smem_u32(smem + V3_SM_Z + pw * 2 * V3_ZSTAGE_W);   // [2 stages][4096]
    const unsigned char* zgen = smem + V3_SM_Z + pw * 2 * V3_ZSTAGE_W;
    const uint32_t zbar = bar_kfull + 32 + 16 * pw;   // + 8 * stage
    const uint4* wbf = reinterpret_cast<const uint4*>(smem + V3_SM_WB) + lane;
    auto issue_z = [&](int jt) {                       // one lane: 4 boxes = (2 rows) x (2 channel halves)
      const int st = jt & 1;
      mbar_arrive_expect_tx(zbar + 8 * st, V3_ZSTAGE_W);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        tma_load_3d(zbase + st * V3_ZSTAGE_W + q * 1024, &args.tm_z, 32 * (q & 1), jt * V2_TK,
                    (int)rowb + i0 + r0 + (q >> 1), zbar + 8 * st);
    };
    float acc[2][4][4];   // o_pair_raw^T: [row][m-tile of 16 channels][C fragment: (ch g | g+8) x (heads 2t, 2t+1)]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { acc[r][mt][0] = acc[r][mt][1] = acc[r][mt][2] = acc[r][mt][3] = 0.f; }
    if (lane == 0) issue_z(0);
    for (int jt = 0; jt < JT; ++jt) {
      const int par = jt & 1;
      __syncwarp();                                    // every lane is done with the other stage (tile jt - 1)
      if (lane == 0 && jt + 1 < JT) issue_z(jt + 1);
      mbar_wait(zbar + 8 * par, (jt >> 1) & 1);        // z tile jt (this warp's rows) has landed
      const unsigned char* zt = zgen + par * V3_ZSTAGE_W;
      // ---- pair bias for rows r0, r0+1 x 8 keys, all heads -> sbias[par]
      {
        float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const int c = ks * 16 + 2 * t;
          const float2 x0 = *reinterpret_cast<const float2*>(zt + v3_zoff(0, g, c));       // pair (row r0,   key g)
          const float2 x1 = *reinterpret_cast<const float2*>(zt + v3_zoff(1, g, c));       // pair (row r0+1, key g)
          const float2 x2 = *reinterpret_cast<const float2*>(zt + v3_zoff(0, g, c + 8));
          const float2 x3 = *reinterpret_cast<const float2*>(zt + v3_zoff(1, g, c + 8));
          uint32_t ah[4], al[4];
          split_pair(x0.x, x0.y, ah[0], al[0]);
          split_pair(x1.x, x1.y, ah[1], al[1]);
          split_pair(x2.x, x2.y, ah[2], al[2]);
          split_pair(x3.x, x3.y, ah[3], al[3]);
          const uint4 w = wbf[ks * 32];
          mma16816(c0, al, w.x, w.y);
          mma16816(c1, ah, w.x, w.y);
          mma16816(c0, ah, w.z, w.w);
        }
        float* bs = sbias + par * (H * V2_TQ * V2_BP);
        bs[((2 * t) * V2_TQ + r0) * V2_BP + g] = c0[0] + c1[0];
        bs[((2 * t + 1) * V2_TQ + r0) * V2_BP + g] = c0[1] + c1[1];
        bs[((2 * t) * V2_TQ + r0 + 1) * V2_BP + g] = c0[2] + c1[2];
        bs[((2 * t + 1) * V2_TQ + r0 + 1) * V2_BP + g] = c0[3] + c1[3];
      }
      named_arrive(1 + par, V3_THREADS);
      // ---- o_pair_raw^T[c, h] = alpha_h * old + sum_j z[j, c] P[h, j] once the head warps have published P / alpha:
      //      m16n8k8 with M = 16 channels, N = 8 heads, K = the 8 keys of the tile (3xFP16)
      named_sync(3 + par, V3_THREADS);
      {
        const float* Pr = sP + par * (V2_TQ * V2_PP);
        const float* al = salpha + par * (H * V2_TQ);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float a0 = al[(2 * t) * V2_TQ + r0 + r], a1 = al[(2 * t + 1) * V2_TQ + r0 + r];
          uint32_t bh, bl;                              // B = P^T: (keys 2t, 2t+1; head g)
          split_pair(Pr[(r0 + r) * V2_PP + (2 * t) * 8 + g], Pr[(r0 + r) * V2_PP + (2 * t + 1) * 8 + g], bh, bl);
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            const int cA = 16 * mt + g, cB = cA + 8;
            const float zAA = *reinterpret_cast<const float*>(zt + v3_zoff(r, 2 * t, cA));
            const float zBA = *reinterpret_cast<const float*>(zt + v3_zoff(r, 2 * t + 1, cA));
            const float zAB = *reinterpret_cast<const float*>(zt + v3_zoff(r, 2 * t, cB));
            const float zBB = *reinterpret_cast<const float*>(zt + v3_zoff(r, 2 * t + 1, cB));
            uint32_t ah0, al0, ah1, al1;
            split_pair(zAA, zBA, ah0, al0);             // a0: (channel cA; keys 2t, 2t+1)
            split_pair(zAB, zBB, ah1, al1);             // a1: (channel cB; keys 2t, 2t+1)
            acc[r][mt][0] *= a0; acc[r][mt][1] *= a1; acc[r][mt][2] *= a0; acc[r][mt][3] *= a1;
            mma1688(acc[r][mt], al0, al1, bh);
            mma1688(acc[r][mt], ah0, ah1, bl);
            mma1688(acc[r][mt], ah0, ah1, bh);
          }
        }
      }
    }
    named_sync(5, V3_THREADS);                         // every warp has left the tile loop: the blob region is free
    float* sop = reinterpret_cast<float*>(smem + V3_SM_BLOB);   // [16 i][8 h][64]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float* o = sop + ((r0 + r) * H + 2 * t) * CZ + 16 * mt + g;
        o[0] = acc[r][mt][0]; o[CZ] = acc[r][mt][1];
        o[8] = acc[r][mt][2]; o[CZ + 8] = acc[r][mt][3];
      }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;\n");
  }
  __syncthreads();

  // ---- common epilogue (all 16 warps)
  const float* spt = reinterpret_cast<const float*>(smem + V3_SM_Z);
  const float* sop = reinterpret_cast<const float*>(smem + V3_SM_BLOB);
  // o_pt: global -> local frame, norms (ipa_pytorch.py:455-460)
  for (int idx = tid; idx < H * V2_TQ * PV; idx += V3_THREADS) {
    const int pnt = idx % PV, i = (idx / PV) % V2_TQ, hh = idx / (PV * V2_TQ);
    if (i0 + i >= L) continue;
    const float* R = a.rot + (rowb + i0 + i) * 9;
    const float* tr = a.trans + (rowb + i0 + i) * 3;
    const float* s = spt + (hh * V2_TQ + i) * 40 + pnt * 3;
    const float gx = s[0] - tr[0], gy = s[1] - tr[1], gz = s[2] - tr[2];
    const float lx = R[0] * gx + R[3] * gy + R[6] * gz;
    const float ly = R[1] * gx + R[4] * gy + R[7] * gz;
    const float lz = R[2] * gx + R[5] * gy + R[8] * gz;
    float* f = a.feats + (rowb + i0 + i) * NFEAT + 1024;
    f[0 * 96 + hh * PV + pnt] = lx;
    f[1 * 96 + hh * PV + pnt] = ly;
    f[2 * 96 + hh * PV + pnt] = lz;
    f[3 * 96 + hh * PV + pnt] = sqrtf(lx * lx + ly * ly + lz * lz + 1e-8f);
  }
  // o_pair: down_z on the normalised a-weighted pair row (ipa_pytorch.py:469-473); W_dz staged in shared memory
  float* swz = reinterpret_cast<float*>(smem + V3_SM_QLO);      // [16 d][68]
  for (int idx = tid; idx < 16 * CZ; idx += V3_THREADS) swz[(idx >> 6) * V2_ZP + (idx & 63)] = a.w_dz[idx];
  __syncthreads();
  {
    const int pairidx = tid >> 2, d0 = (tid & 3) * 4;   // pairidx = i * 8 + head; 4 of the 16 outputs per thread
    const int i = pairidx >> 3, hh = pairidx & 7;
    if (i0 + i < L) {
      const float* src = sop + (i * H + hh) * CZ;
      float acc4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int c = 0; c < CZ; c += 4) {
        const float4 x = *reinterpret_cast<const float4*>(src + c);
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float4 w = *reinterpret_cast<const float4*>(swz + (d0 + d) * V2_ZP + c);
          acc4[d] = fmaf(w.x, x.x, acc4[d]);
          acc4[d] = fmaf(w.y, x.y, acc4[d]);
          acc4[d] = fmaf(w.z, x.z, acc4[d]);
          acc4[d] = fmaf(w.w, x.w, acc4[d]);
        }
      }
      const float inv = sl[hh * V2_TQ + i];
      float4 o;
      o.x = acc4[0] * inv + a.b_dz[d0 + 0]; o.y = acc4[1] * inv + a.b_dz[d0 + 1];
      o.z = acc4[2] * inv + a.b_dz[d0 + 2]; o.w = acc4[3] * inv + a.b_dz[d0 + 3];
      *reinterpret_cast<float4*>(a.feats + (rowb + i0 + i) * NFEAT + 1024 + 384 + hh * 16 + d0) = o;
    }
  }
}

int launch_ipa_attention_v3(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (a.B == 0 || a.L == 0) return PF_OK;
  PF_REQUIRE(workspace && workspace_bytes >= ipa_v2_workspace_bytes(a.B, a.L), PF_ERR_WORKSPACE_TOO_SMALL);
  const int JT = (a.L + V2_TK - 1) / V2_TK, IT = (a.L + V2_TQ - 1) / V2_TQ;
  unsigned char* blobs = static_cast<unsigned char*>(workspace);
  uint4* Qp = reinterpret_cast<uint4*>(blobs + (size_t)a.B * JT * V2_BLOB);
  IpaPack2Args pa{a.proj, a.pts, a.head_w, a.mask, blobs, Qp, a.B, a.L, JT, IT};
  const size_t total = (size_t)a.B * JT * H * V2_KS * 32 + (size_t)a.B * JT * H * V2_VNT * 32 +
                       (size_t)a.B * JT * (H + 1) * V2_TK + (size_t)a.B * H * IT * V2_KS * 32;
  ipa_pack2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pa);
  PF_CHECK_LAUNCH();
  Ipa3Args args;
  PF_TRY(encode_z_map(&args.tm_z, a.z, a.B, a.L));
  args.p = Ipa2Args{a, blobs, Qp, JT, IT};
  profile_begin(0, st);
  ipa_attention_v3_kernel<<<dim3(IT, a.B), V3_THREADS, V3_SMEM, st>>>(args);
  profile_end(0, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_ipa_attention_v2(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (a.B == 0 || a.L == 0) return PF_OK;
  PF_REQUIRE(workspace && workspace_bytes >= ipa_v2_workspace_bytes(a.B, a.L), PF_ERR_WORKSPACE_TOO_SMALL);
  const int JT = (a.L + V2_TK - 1) / V2_TK, IT = (a.L + V2_TQ - 1) / V2_TQ;
  unsigned char* blobs = static_cast<unsigned char*>(workspace);
  uint4* Qp = reinterpret_cast<uint4*>(blobs + (size_t)a.B * JT * V2_BLOB);
  IpaPack2Args pa{a.proj, a.pts, a.head_w, a.mask, blobs, Qp, a.B, a.L, JT, IT};
  const size_t total = (size_t)a.B * JT * H * V2_KS * 32 + (size_t)a.B * JT * H * V2_VNT * 32 +
                       (size_t)a.B * JT * (H + 1) * V2_TK + (size_t)a.B * H * IT * V2_KS * 32;
  ipa_pack2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pa);
  PF_CHECK_LAUNCH();
  Ipa2Args p{a, blobs, Qp, JT, IT};
  profile_begin(0, st);
  ipa_attention_v2_kernel<<<dim3(IT, a.B), 256, V2_SMEM, st>>>(p);
  profile_end(0, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

void ipa_v2_kernels_init() {
  cudaFuncSetAttribute(ipa_attention_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V2_SMEM);
  cudaFuncSetAttribute(ipa_attention_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V3_SMEM);
}

}  // namespace pf
