// K3, variant 1 ("ipa_impl" = 1): tensor-core fused invariant-point attention.
//
// CTA = 16 query rows x all 8 heads of one complex; keys are streamed in tiles of 16 with an online
// softmax, so nothing of size L^2 x H ever leaves the SM and the pair tensor z is read from HBM exactly once.
//
//   warp roles change per phase (all 8 warps take part in every phase):
//     pair-major  (warp w owns query rows 2w, 2w+1):   pair bias  W_b z   (3xFP16 mma, z tile in smem)
//     head-major  (warp h owns head h):                 S = Q K^T  (3xFP16 mma, K fragments pre-packed)
//                                                        + bias + point distances (fp32 FMA, direct differences)
//                                                        online softmax, O += P [V | v_pts] (3xFP16 mma)
//     pair-major:                                        o_pair_raw[i,h,:] += sum_j P[h,i,j] z[i,j,:]  (fp32 FMA)
//   z tiles (16 x 16 pairs x 64 ch, padded rows) and the key points arrive through a 2-stage cp.async ring.
//
// The q/k/v operands are converted once per call by ipa_pack_kernel into fp16 hi/lo mma fragments
// (pf_split.cuh) laid out so that each lane fetches its fragment with one 16-byte load.
#include "pf_common.cuh"
#include "pf_split.cuh"

namespace pf {

constexpr int TQ = 16;                 // query rows per CTA
constexpr int TKEY = 16;               // keys per tile
constexpr int ZP = 68;                 // padded z row (floats)
constexpr int VNT = 21;                // n-tiles of [v(128) | v_pts(36) | pad(4)]
constexpr int PP = 17;                 // pitch of the P tile rows
constexpr int BP = 20;                 // pitch of the bias tile rows

// ---- packed operand sizes (uint4 units)
__host__ __device__ inline size_t kp_tile_u4() { return 8 * 2 * 32; }      // per (b,h,jt): 8 k-steps x 2 n-tiles x 32 lanes
__host__ __device__ inline size_t vp_tile_u4() { return VNT * 32; }        // per (b,h,jt)
__host__ __device__ inline size_t qp_tile_u4() { return 8 * 32 * 2; }      // per (b,h,it): 8 k-steps x 32 lanes x (hi,lo)

struct IpaPackArgs {
  const float* proj; const float* pts;
  uint4* Kp; uint4* Vp; uint4* Qp;
  int B, L, JT, IT;
};

__global__ void ipa_pack_kernel(IpaPackArgs a) {
  const size_t nK = (size_t)a.B * H * a.JT * kp_tile_u4();
  const size_t nV = (size_t)a.B * H * a.JT * vp_tile_u4();
  const size_t nQ = (size_t)a.B * H * a.IT * (8 * 32);
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const int L = a.L;
  if (idx < nK) {
    const int lane = idx & 31, nt = (idx >> 5) & 1, ks = (idx >> 6) & 7;
    const size_t rest = idx >> 9;
    const int jt = rest % a.JT, h = (rest / a.JT) % H, b = rest / ((size_t)a.JT * H);
    const int g = lane >> 2, t = lane & 3;
    const int j = jt * TKEY + nt * 8 + g;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < L) {
      const float* k = a.proj + ((size_t)b * L + j) * NPROJ + OFF_KV + h * 2 * C + ks * 16 + 2 * t;
      v[0] = k[0]; v[1] = k[1]; v[2] = k[8]; v[3] = k[9];
    }
    uint4 o;
    split_pair(v[0], v[1], o.x, o.z);
    split_pair(v[2], v[3], o.y, o.w);
    a.Kp[idx] = o;
  } else if (idx < nK + nV) {
    const size_t r = idx - nK;
    const int lane = r & 31;
    const int nt = (r >> 5) % VNT;
    const size_t rest = (r >> 5) / VNT;
    const int jt = rest % a.JT, h = (rest / a.JT) % H, b = rest / ((size_t)a.JT * H);
    const int g = lane >> 2, t = lane & 3;
    const int n = nt * 8 + g;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = jt * TKEY + 2 * t + (e & 1) + (e >> 1) * 8;
      float x = 0.f;
      if (j < L) {
        const size_t row = (size_t)b * L + j;
        if (n < C) x = a.proj[row * NPROJ + OFF_KV + h * 2 * C + C + n];
        else if (n < C + PV * 3) x = a.pts[(row * H + h) * (NPT * 3) + 2 * PQ * 3 + (n - C)];
      }
      v[e] = x;
    }
    uint4 o;
    split_pair(v[0], v[1], o.x, o.z);
    split_pair(v[2], v[3], o.y, o.w);
    a.Vp[r] = o;
  } else if (idx < nK + nV + nQ) {
    const size_t r = idx - nK - nV;
    const int lane = r & 31, ks = (r >> 5) & 7;
    const size_t rest = r >> 8;
    const int it = rest % a.IT, h = (rest / a.IT) % H, b = rest / ((size_t)a.IT * H);
    const int g = lane >> 2, t = lane & 3;
    const float sc = 0.05103103630798288f;  // sqrt(1/(3*128)) folded into q
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = it * TQ + g + half * 8;
      if (i < L) {
        const float* q = a.proj + ((size_t)b * L + i) * NPROJ + OFF_Q + h * C + ks * 16 + 2 * t;
        v[half * 2 + 0] = q[0] * sc; v[half * 2 + 1] = q[1] * sc;       // a0 (row g) / a1 (row g+8): k = 2t, 2t+1
        v[4 + half * 2 + 0] = q[8] * sc; v[4 + half * 2 + 1] = q[9] * sc;  // a2 / a3: k = 2t+8, 2t+9
      }
    }
    uint4 hi, lo;
    split_pair(v[0], v[1], hi.x, lo.x);
    split_pair(v[2], v[3], hi.y, lo.y);
    split_pair(v[4], v[5], hi.z, lo.z);
    split_pair(v[6], v[7], hi.w, lo.w);
    a.Qp[r * 2] = hi;
    a.Qp[r * 2 + 1] = lo;
  }
}

// ---- shared memory carve-up (floats unless noted)
constexpr int SM_Z_STAGE = TQ * TKEY * ZP;          // 17408 floats
constexpr int SM_KPT_STAGE = H * TKEY * PQ * 3;      // 3072
constexpr int SM_QPT = H * TQ * PQ * 3;              // 3072
constexpr int SM_BIAS = H * TQ * BP;                 // 2560
constexpr int SM_P = TQ * H * PP;                    // 2176
constexpr int SM_ALPHA = H * TQ;                     // 128
constexpr int SM_OPAIR = TQ * H * CZ;                // 8192
constexpr int SM_MJ = 2 * TKEY;                      // 32
constexpr int SM_TOTAL_FLOATS = 2 * SM_Z_STAGE + 2 * SM_KPT_STAGE + SM_QPT + SM_BIAS + SM_P + 2 * SM_ALPHA + SM_OPAIR + SM_MJ;
constexpr size_t IPA1_SMEM = (size_t)SM_TOTAL_FLOATS * sizeof(float);

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}

struct Ipa1Args {
  IpaArgs a;
  const uint4* Kp; const uint4* Vp; const uint4* Qp;
  int JT, IT;
};

__global__ void __launch_bounds__(256, 1) ipa_attention_v1_kernel(Ipa1Args p) {
  extern __shared__ __align__(16) float smem[];
  float* zs = smem;                                   // [2][16 i][16 j][68]
  float* kpt = zs + 2 * SM_Z_STAGE;                   // [2][8 h][16 j][24]
  float* qpt = kpt + 2 * SM_KPT_STAGE;                // [8 h][16 i][24]
  float* sbias = qpt + SM_QPT;                        // [8 h][16 i][20]
  float* sP = sbias + SM_BIAS;                        // [16 i][8 h][17]
  float* salpha = sP + SM_P;                          // [8 h][16 i]
  float* sl = salpha + SM_ALPHA;                      // [8 h][16 i]   final row sums
  float* sop = sl + SM_ALPHA;                         // [16 i][8 h][64]
  float* smj = sop + SM_OPAIR;                        // [2][16]
  const IpaArgs& a = p.a;
  const int L = a.L;
  const int b = blockIdx.y, it = blockIdx.x, i0 = it * TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const size_t rowb = (size_t)b * L;
  const int h = warp;                                  // head-major role
  const float sc_b = 0.5773502691896257f;              // sqrt(1/3)

  // ---- one-time loads
  for (int idx = tid; idx < SM_QPT; idx += 256) {
    const int e = idx % 24, i = (idx / 24) % TQ, hh = idx / (24 * TQ);
    qpt[idx] = (i0 + i < L) ? a.pts[((rowb + i0 + i) * H + hh) * (NPT * 3) + e] : 0.f;
  }
  for (int idx = tid; idx < SM_OPAIR; idx += 256) sop[idx] = 0.f;
  // W_b as B fragments (pair-bias mma): rows n = head (8), k = channel; sqrt(1/3) folded in
  uint32_t wbh[4][2], wbl[4][2];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const float* w = a.w_b + g * CZ + ks * 16 + 2 * t;
    split_pair(w[0] * sc_b, w[1] * sc_b, wbh[ks][0], wbl[ks][0]);
    split_pair(w[8] * sc_b, w[9] * sc_b, wbh[ks][1], wbl[ks][1]);
  }
  // Q fragments of head h (scaled, hi/lo)
  uint4 qh[8], ql[8];
  {
    const uint4* qp = p.Qp + (((size_t)b * H + h) * p.IT + it) * qp_tile_u4();
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qh[ks] = qp[(ks * 32 + lane) * 2];
      ql[ks] = qp[(ks * 32 + lane) * 2 + 1];
    }
  }
  const float hw = a.head_w[h];
  const float bbias = sc_b * a.b_b[h];
  const float mi_lo = (i0 + g < L) ? a.mask[rowb + i0 + g] : 0.f;
  const float mi_hi = (i0 + g + 8 < L) ? a.mask[rowb + i0 + g + 8] : 0.f;

  float O[VNT][4];
#pragma unroll
  for (int n = 0; n < VNT; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  auto issue_tile = [&](int jt) {
    const int st = jt & 1;
    const int j0 = jt * TKEY;
    float* zd = zs + st * SM_Z_STAGE;
    for (int c = tid; c < TQ * TKEY * 16; c += 256) {          // 16-byte chunks: pair (i,j) x 16 chunks
      const int ch = c & 15, pr = c >> 4, j = pr & 15, i = pr >> 4;
      const bool ok = (i0 + i < L) && (j0 + j < L);
      const float* src = a.z + (((rowb + (ok ? i0 + i : 0)) * L + (ok ? j0 + j : 0)) * CZ) + ch * 4;
      cp_async16_zfill(zd + pr * ZP + ch * 4, src, ok);
    }
    float* kd = kpt + st * SM_KPT_STAGE;
    for (int c = tid; c < H * TKEY * 6; c += 256) {             // 96-byte (24-float) key-point rows
      const int ch = c % 6, j = (c / 6) % TKEY, hh = c / (6 * TKEY);
      const bool ok = (j0 + j < L);
      const float* src = a.pts + ((rowb + (ok ? j0 + j : 0)) * H + hh) * (NPT * 3) + PQ * 3 + ch * 4;
      cp_async16_zfill(kd + (hh * TKEY + j) * 24 + ch * 4, src, ok);
    }
    if (tid < TKEY) smj[st * TKEY + tid] = (j0 + tid < L) ? a.mask[rowb + j0 + tid] : 0.f;
    asm volatile("cp.async.commit_group;\n" ::);
  };

  issue_tile(0);
  for (int jt = 0; jt < p.JT; ++jt) {
    const int st = jt & 1;
    const int j0 = jt * TKEY;
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();                                   // tile jt visible; everyone finished tile jt-1
    if (jt + 1 < p.JT) issue_tile(jt + 1);
    const float* zt = zs + st * SM_Z_STAGE;

    // ================= phase 1 (pair-major): pair bias for rows 2w, 2w+1, all heads
    {
      float acc[2][2][4];                              // [m-tile][small terms | hi*hi]: 4 independent MMA chains
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int q = 0; q < 2; ++q) { acc[mt][q][0] = acc[mt][q][1] = acc[mt][q][2] = acc[mt][q][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {               // m-tile = one query row (i = 2w + mt), 16 keys
          const int i = 2 * warp + mt;
          const float* zr_lo = zt + (i * TKEY + g) * ZP;       // pair (i, j = g)
          const float* zr_hi = zt + (i * TKEY + g + 8) * ZP;   // pair (i, j = g + 8)
          const int c = ks * 16 + 2 * t;
          const float2 x0 = *reinterpret_cast<const float2*>(zr_lo + c);
          const float2 x1 = *reinterpret_cast<const float2*>(zr_hi + c);
          const float2 x2 = *reinterpret_cast<const float2*>(zr_lo + c + 8);
          const float2 x3 = *reinterpret_cast<const float2*>(zr_hi + c + 8);
          split_pair(x0.x, x0.y, ah[mt][0], al[mt][0]);
          split_pair(x1.x, x1.y, ah[mt][1], al[mt][1]);
          split_pair(x2.x, x2.y, ah[mt][2], al[mt][2]);
          split_pair(x3.x, x3.y, ah[mt][3], al[mt][3]);
        }
        mma16816(acc[0][0], al[0], wbh[ks][0], wbh[ks][1]);
        mma16816(acc[1][0], al[1], wbh[ks][0], wbh[ks][1]);
        mma16816(acc[0][1], ah[0], wbh[ks][0], wbh[ks][1]);
        mma16816(acc[1][1], ah[1], wbh[ks][0], wbh[ks][1]);
        mma16816(acc[0][0], ah[0], wbl[ks][0], wbl[ks][1]);
        mma16816(acc[1][0], ah[1], wbl[ks][0], wbl[ks][1]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int i = 2 * warp + mt;
        // C: (pair row g -> key j = g, head 2t,2t+1), (row g+8 -> key g+8)
        sbias[((2 * t) * TQ + i) * BP + g] = acc[mt][0][0] + acc[mt][1][0];
        sbias[((2 * t + 1) * TQ + i) * BP + g] = acc[mt][0][1] + acc[mt][1][1];
        sbias[((2 * t) * TQ + i) * BP + g + 8] = acc[mt][0][2] + acc[mt][1][2];
        sbias[((2 * t + 1) * TQ + i) * BP + g + 8] = acc[mt][0][3] + acc[mt][1][3];
      }
    }
    __syncthreads();                                   // (A) bias tile complete

    // ================= phase 2 (head-major): S = Q K^T for head h
    float S[2][4];
    {
      float Sa[2][4], Sb[2][4];                        // [n-tile]: small terms / hi*hi -> 4 independent chains
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        Sa[nt][0] = Sa[nt][1] = Sa[nt][2] = Sa[nt][3] = 0.f;
        Sb[nt][0] = Sb[nt][1] = Sb[nt][2] = Sb[nt][3] = 0.f;
      }
      const uint4* kp = p.Kp + (((size_t)b * H + h) * p.JT + jt) * kp_tile_u4();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint4 k0 = kp[(ks * 2 + 0) * 32 + lane];
        const uint4 k1 = kp[(ks * 2 + 1) * 32 + lane];
        const uint32_t ah[4] = {qh[ks].x, qh[ks].y, qh[ks].z, qh[ks].w};
        const uint32_t al[4] = {ql[ks].x, ql[ks].y, ql[ks].z, ql[ks].w};
        mma16816(Sa[0], al, k0.x, k0.y);
        mma16816(Sa[1], al, k1.x, k1.y);
        mma16816(Sb[0], ah, k0.x, k0.y);
        mma16816(Sb[1], ah, k1.x, k1.y);
        mma16816(Sa[0], ah, k0.z, k0.w);
        mma16816(Sa[1], ah, k1.z, k1.w);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) S[nt][e] = Sa[nt][e] + Sb[nt][e];
    }
    // ================= phase 3: logits, online softmax
    {
      const float* kp_s = kpt + st * SM_KPT_STAGE + h * TKEY * 24;
      const float* qp_lo = qpt + (h * TQ + g) * 24;
      const float* qp_hi = qpt + (h * TQ + g + 8) * 24;
      float d2[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { d2[nt][0] = d2[nt][1] = d2[nt][2] = d2[nt][3] = 0.f; }
#pragma unroll
      for (int e4 = 0; e4 < 6; ++e4) {
        const float4 ql4 = *reinterpret_cast<const float4*>(qp_lo + e4 * 4);
        const float4 qh4 = *reinterpret_cast<const float4*>(qp_hi + e4 * 4);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4 k4 = *reinterpret_cast<const float4*>(kp_s + (nt * 8 + 2 * t + e) * 24 + e4 * 4);
            float dx;
            dx = ql4.x - k4.x; d2[nt][e] = fmaf(dx, dx, d2[nt][e]);
            dx = ql4.y - k4.y; d2[nt][e] = fmaf(dx, dx, d2[nt][e]);
            dx = ql4.z - k4.z; d2[nt][e] = fmaf(dx, dx, d2[nt][e]);
            dx = ql4.w - k4.w; d2[nt][e] = fmaf(dx, dx, d2[nt][e]);
            dx = qh4.x - k4.x; d2[nt][2 + e] = fmaf(dx, dx, d2[nt][2 + e]);
            dx = qh4.y - k4.y; d2[nt][2 + e] = fmaf(dx, dx, d2[nt][2 + e]);
            dx = qh4.z - k4.z; d2[nt][2 + e] = fmaf(dx, dx, d2[nt][2 + e]);
            dx = qh4.w - k4.w; d2[nt][2 + e] = fmaf(dx, dx, d2[nt][2 + e]);
          }
      }
      float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * t + e;
          const bool valid = (j0 + j < L);
          const float mj = smj[st * TKEY + j];
          const float b_lo = sbias[(h * TQ + g) * BP + j], b_hi = sbias[(h * TQ + g + 8) * BP + j];
          float x_lo = S[nt][e] + b_lo + bbias - 0.5f * hw * d2[nt][e] + 1e5f * (mi_lo * mj - 1.f);
          float x_hi = S[nt][2 + e] + b_hi + bbias - 0.5f * hw * d2[nt][2 + e] + 1e5f * (mi_hi * mj - 1.f);
          x_lo = valid ? x_lo : -INFINITY;
          x_hi = valid ? x_hi : -INFINITY;
          S[nt][e] = x_lo; S[nt][2 + e] = x_hi;
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
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float p_lo = expf(S[nt][e] - mn_lo), p_hi = expf(S[nt][2 + e] - mn_hi);
          S[nt][e] = p_lo; S[nt][2 + e] = p_hi;
          ps_lo += p_lo; ps_hi += p_hi;
          const int j = nt * 8 + 2 * t + e;
          sP[(g * H + h) * PP + j] = p_lo;
          sP[((g + 8) * H + h) * PP + j] = p_hi;
        }
      l_lo = l_lo * al_lo + ps_lo;
      l_hi = l_hi * al_hi + ps_hi;
      if (t == 0) { salpha[h * TQ + g] = al_lo; salpha[h * TQ + g + 8] = al_hi; }
#pragma unroll
      for (int n = 0; n < VNT; ++n) { O[n][0] *= al_lo; O[n][1] *= al_lo; O[n][2] *= al_hi; O[n][3] *= al_hi; }
    }
    // ================= phase 4: O += P [V | v_pts]
    {
      uint32_t ph[4], pl[4];
      split_pair(S[0][0], S[0][1], ph[0], pl[0]);      // a0: row g,   keys 2t,2t+1
      split_pair(S[0][2], S[0][3], ph[1], pl[1]);      // a1: row g+8
      split_pair(S[1][0], S[1][1], ph[2], pl[2]);      // a2: row g,   keys 2t+8,2t+9
      split_pair(S[1][2], S[1][3], ph[3], pl[3]);      // a3: row g+8
      const uint4* vp = p.Vp + (((size_t)b * H + h) * p.JT + jt) * vp_tile_u4();
#pragma unroll
      for (int n0 = 0; n0 < VNT; n0 += 7) {            // 7 independent accumulator chains in flight
        uint4 v[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) v[q] = vp[(n0 + q) * 32 + lane];
#pragma unroll
        for (int q = 0; q < 7; ++q) mma16816(O[n0 + q], pl, v[q].x, v[q].y);
#pragma unroll
        for (int q = 0; q < 7; ++q) mma16816(O[n0 + q], ph, v[q].z, v[q].w);
#pragma unroll
        for (int q = 0; q < 7; ++q) mma16816(O[n0 + q], ph, v[q].x, v[q].y);
      }
    }
    __syncthreads();                                   // (B) P tile and alpha complete

    // ================= phase 5 (pair-major): o_pair_raw[i, h, :] = alpha * old + sum_j P z
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int i = 2 * warp + mt;
      float acc[H][2];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) {
        const float al = salpha[hh * TQ + i];
        const float2 o = *reinterpret_cast<const float2*>(sop + (i * H + hh) * CZ + 2 * lane);
        acc[hh][0] = o.x * al; acc[hh][1] = o.y * al;
      }
#pragma unroll 4
      for (int j = 0; j < TKEY; ++j) {
        const float2 zv = *reinterpret_cast<const float2*>(zt + (i * TKEY + j) * ZP + 2 * lane);
#pragma unroll
        for (int hh = 0; hh < H; ++hh) {
          const float pv = sP[(i * H + hh) * PP + j];
          acc[hh][0] = fmaf(pv, zv.x, acc[hh][0]);
          acc[hh][1] = fmaf(pv, zv.y, acc[hh][1]);
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
  if (t == 0) { sl[h * TQ + g] = il_lo; sl[h * TQ + g + 8] = il_hi; }
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
    for (int n = 16; n < VNT; ++n) {
      const int c = (n - 16) * 8 + 2 * t;
      spt[(h * TQ + g) * 40 + c] = O[n][0] * il_lo;
      spt[(h * TQ + g) * 40 + c + 1] = O[n][1] * il_lo;
      spt[(h * TQ + g + 8) * 40 + c] = O[n][2] * il_hi;
      spt[(h * TQ + g + 8) * 40 + c + 1] = O[n][3] * il_hi;
    }
  }
  __syncthreads();
  // o_pt: global -> local frame, norms (ipa_pytorch.py:455-460)
  for (int idx = tid; idx < H * TQ * PV; idx += 256) {
    const int pnt = idx % PV, i = (idx / PV) % TQ, hh = idx / (PV * TQ);
    if (i0 + i >= L) continue;
    const float* R = a.rot + (rowb + i0 + i) * 9;
    const float* tr = a.trans + (rowb + i0 + i) * 3;
    const float* s = spt + (hh * TQ + i) * 40 + pnt * 3;
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
  // o_pair: down_z on the normalised a-weighted pair row (ipa_pytorch.py:469-473).  W_dz is staged in shared
  // memory (the o_pt scratch of the loop above is dead after the barrier); thread = (row i, head, half of the
  // 16 outputs): 8 independent dot products of length 64, operands read as 16-byte vectors.
  __syncthreads();
  float* swz = zs;                                      // [16 d][68]
  for (int idx = tid; idx < 16 * CZ; idx += 256) swz[(idx >> 6) * ZP + (idx & 63)] = a.w_dz[idx];
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
          const float4 w = *reinterpret_cast<const float4*>(swz + (d0 + d) * ZP + c);
          acc[d] = fmaf(w.x, x.x, acc[d]);
          acc[d] = fmaf(w.y, x.y, acc[d]);
          acc[d] = fmaf(w.z, x.z, acc[d]);
          acc[d] = fmaf(w.w, x.w, acc[d]);
        }
      }
      const float inv = sl[hh * TQ + i];
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

size_t ipa_workspace_bytes(int B, int L) {
  const size_t JT = (L + TKEY - 1) / TKEY, IT = (L + TQ - 1) / TQ;
  const size_t u4 = (size_t)B * H * (JT * (kp_tile_u4() + vp_tile_u4()) + IT * qp_tile_u4());
  const size_t v1 = u4 * sizeof(uint4) + 1024, v2 = ipa_v2_workspace_bytes(B, L);
  return v1 > v2 ? v1 : v2;   // one buffer serves whichever variant is selected
}

int launch_ipa_attention_v1(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (a.B == 0 || a.L == 0) return PF_OK;
  PF_REQUIRE(workspace && workspace_bytes >= ipa_workspace_bytes(a.B, a.L), PF_ERR_WORKSPACE_TOO_SMALL);
  const int JT = (a.L + TKEY - 1) / TKEY, IT = (a.L + TQ - 1) / TQ;
  uint4* Kp = reinterpret_cast<uint4*>(workspace);
  uint4* Vp = Kp + (size_t)a.B * H * JT * kp_tile_u4();
  uint4* Qp = Vp + (size_t)a.B * H * JT * vp_tile_u4();
  IpaPackArgs pa{a.proj, a.pts, Kp, Vp, Qp, a.B, a.L, JT, IT};
  const size_t total = (size_t)a.B * H * (JT * (kp_tile_u4() + vp_tile_u4()) + IT * (8 * 32));
  ipa_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(pa);
  PF_CHECK_LAUNCH();
  Ipa1Args p{a, Kp, Vp, Qp, JT, IT};
  profile_begin(0, st);
  ipa_attention_v1_kernel<<<dim3(IT, a.B), 256, IPA1_SMEM, st>>>(p);
  profile_end(0, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

void ipa_tc_kernels_init() {
  cudaFuncSetAttribute(ipa_attention_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IPA1_SMEM);
}

}  // namespace pf
