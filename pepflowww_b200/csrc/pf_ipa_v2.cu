// K3: tensor-core fused invariant-point attention (models_con/ipa_pytorch.py:393-473), "ipa_impl" 3 / 4 (default 4).
//
//   * The point-distance term rides on the tensor core, arranged so that nothing of the size |t|^2 is ever rounded.
//     With q_p = t_i + a_ip, k_p = t_j + b_jp (a, b = the ROTATED local points; t = frame translations),
//       sum_p |q_p - k_p|^2 = P |t_i - t_j|^2 + 2 (t_i - t_j).(A_i - B_j) + sum_p |a_ip - b_jp|^2,   A = sum_p a, B = sum_p b.
//     Terms constant along a softmax row drop out; what is left of  -1/2 c_h sum_p |q_p - k_p|^2  is
//       -P/2 c_h |t_i - t_j|^2                         exact fp32 per pair in the head warps (key translations ride in the blob)
//       + c_h (t_i.B_j + A_i.t_j + sum_p a_ip.b_jp)      30 more channels of the scalar Q K^T contraction (128 -> 158 of 160):
//                                                      Q' = [q/sqrt(3C) | c_h t_i | c_h A_i | c_h a_ip],  K' = [k | B_j | t_j | b_jp]
//       - c_h t_j.B_j - 1/2 c_h sum_p |b_jp|^2          per-(key, head) bias computed once by the pack kernel in fp32.
//     The products that reach the accumulator are of size |t| |B| c_h (~75 at 47 A) instead of P |t|^2 c_h (~1800): the
//     expanded form  c_h q.k - 1/2 c_h |k|^2  of round 1 lost 1e-4 on a logit to fp32 cancellation at the bench shape.
//   * Keys are streamed in tiles of 8.  The pre-packed K' and V' fragments of all heads for one tile, the key
//     biases, key mask and key translations form one contiguous blob per (complex, key tile); every head warp bulk-copies
//     the slice of its head (TMA engine, mbarrier completion) while the previous tile is finished; the MMA
//     operands are then conflict-free 16 / 8 byte shared-memory loads instead of L2 round trips.
// CTA = 16 query rows x 8 heads; 8 head warps (Q'K'^T, online softmax, P V) and 8 pair warps (pair bias W_b z on the
// tensor core, o_pair accumulation) - see the kernel below.
#include <cuda.h>   // CUtensorMap

#include "pf_common.cuh"
#include "pf_split.cuh"
#include "pf_umma.cuh"

namespace pf {

using namespace umma;

constexpr int V2_TQ = 16, V2_TK = 8;
constexpr int V2_ZP = 68;            // padded pair row (floats)
constexpr int V2_KS = 10;            // K steps of Q'K'^T: 128 scalar + 30 point-term + 2 zero channels
constexpr int V2_KPW = 6 + PQ * 3;   // 30 point-term channels per head: [B | t | b_p] (keys), [t | A | a_p] (queries)
constexpr int V2_VNT = 21;           // n-tiles of [v(128) | v_pts(36) | pad(4)]
constexpr int V2_BLOB_K = H * V2_KS * 32 * 16;    // 40960: [h][ks][lane] uint4 {hi b0, hi b1, lo b0, lo b1}
constexpr int V2_BLOB_V = H * V2_VNT * 32 * 8;    // 43008: [h][nt][lane] uint2 {hi, lo}
constexpr int V2_BLOB_KB = H * V2_TK * 4;         // 256:   [h][key] fp32  -c_h t_j.B_j - 1/2 c_h sum |b_jp|^2
constexpr int V2_BLOB_M = V2_TK * 4;              // 32:    [key] fp32 residue mask
constexpr int V2_BLOB_T = V2_TK * 4 * 4;          // 128:   [key][4] fp32 frame translation (x, y, z, 0)
constexpr int V2_BLOB = V2_BLOB_K + V2_BLOB_V + V2_BLOB_KB + V2_BLOB_M + V2_BLOB_T;   // 84384 (multiple of 16)
// blob layout: K' | key bias | key mask | key translations | V'
constexpr int V2_OFF_KB = V2_BLOB_K, V2_OFF_M = V2_OFF_KB + V2_BLOB_KB, V2_OFF_T = V2_OFF_M + V2_BLOB_M,
              V2_OFF_V = V2_OFF_T + V2_BLOB_T;
static_assert(V2_OFF_V % 16 == 0 && V2_BLOB % 16 == 0, "V' part / blobs must stay 16-byte aligned");
constexpr int V2_QTILE_U4 = V2_KS * 32 * 2;       // per (b, h, it): [ks][lane]{hi, lo} uint4
constexpr float V2_QSCALE = 0.05103103630798288f; // sqrt(1/(3*128))

// ---- operand packers: K' / V' blobs (ipa_pack4_kernel) and Q' tiles (ipa_packq_kernel) straight from the projection
// and the frames.  The rows a CTA needs are first staged in shared memory with coalesced loads / bulk copies, the local
// points are rotated there (and the value points moved to the global frame, Rigid.apply, rigid_utils.py:1124-1136 - no
// separate ipa_points pass, no points buffer), and the fragments are then cut from shared memory; every HBM access is a
// full line.
constexpr int P3_THREADS = 512;
constexpr int P3_KVP = 2048 + 8;                  // padded k|v row (floats): fragment loads spread over the banks
constexpr int P3_NP = PQ + PV;                    // 20 key / value points per head

struct IpaPack3Args {
  const float* proj; const float* rot; const float* trans; const float* head_w; const float* mask;
  unsigned char* blobs; uint4* Qp;
  int B, L, JT, IT;
};

// Q' tile of 16 query rows [b][h][it][ks][lane]{hi, lo}; shared memory: [16][192] local q points | [16][8 h][30]
// point-term channels (t | A | a_p) | [16][12] frames
constexpr int PQ_SMEM = (V2_TQ * 192 + V2_TQ * H * V2_KPW + V2_TQ * 12) * 4;
__global__ void __launch_bounds__(P3_THREADS) ipa_packq_kernel(IpaPack3Args a) {
  extern __shared__ __align__(16) float p3[];
  const int L = a.L, tid = threadIdx.x;
  const int qi = blockIdx.x;
  const int b = qi / a.IT, it = qi - b * a.IT, i0 = it * V2_TQ;
  const size_t row0 = (size_t)b * L;
  const int nr = min(V2_TQ, L - i0);
  float* s_loc = p3;                                                   // [16][192] local q points [xyz][h][8]
  float* s_qp = p3 + V2_TQ * 192;                                      // [16][8 h][30]
  float* s_fr = s_qp + V2_TQ * H * V2_KPW;
  for (int i = tid; i < V2_TQ * 48; i += P3_THREADS) {
    const int k = i / 48, q = i - k * 48;
    const float4 v = k < nr ? __ldg(reinterpret_cast<const float4*>(a.proj + (row0 + i0 + k) * NPROJ + OFF_QP) + q)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(s_loc + k * 192 + 4 * q) = v;
  }
  if (tid < V2_TQ * 12) {
    const int k = tid / 12, e = tid - k * 12;
    s_fr[tid] = k < nr ? (e < 9 ? a.rot[(row0 + i0 + k) * 9 + e] : a.trans[(row0 + i0 + k) * 3 + e - 9]) : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < V2_TQ * H * PQ; i += P3_THREADS) {             // a_ip = R_i l_ip (rotation only)
    const int k = i / (H * PQ), r = i - k * (H * PQ), h = r / PQ, pnt = r - h * PQ;
    const float* lp = s_loc + k * 192 + h * PQ + pnt;
    const float lx = lp[0], ly = lp[H * PQ], lz = lp[2 * H * PQ];
    const float* R = s_fr + k * 12;
    float* dst = s_qp + (k * H + h) * V2_KPW + 6 + pnt * 3;
    const bool on = k < nr;
    dst[0] = on ? R[0] * lx + R[1] * ly + R[2] * lz : 0.f;
    dst[1] = on ? R[3] * lx + R[4] * ly + R[5] * lz : 0.f;
    dst[2] = on ? R[6] * lx + R[7] * ly + R[8] * lz : 0.f;
  }
  __syncthreads();
  if (tid < V2_TQ * H * 3) {                                           // t_i and A_i = sum_p a_ip
    const int k = tid / (H * 3), r = tid - k * (H * 3), h = r / 3, x = r - h * 3;
    float* row = s_qp + (k * H + h) * V2_KPW;
    float sum = 0.f;
#pragma unroll
    for (int pnt = 0; pnt < PQ; ++pnt) sum += row[6 + pnt * 3 + x];
    row[x] = s_fr[k * 12 + 9 + x];
    row[3 + x] = sum;
  }
  __syncthreads();
  for (int idx = tid; idx < H * V2_KS * 32; idx += P3_THREADS) {
    const int lane = idx & 31, r = idx >> 5, ks = r % V2_KS, h = r / V2_KS;
    const int g = lane >> 2, t = lane & 3, kk = ks * 16 + 2 * t;
    const float ch = a.head_w[h];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int k = g + half * 8;
      if (k >= nr) continue;
      if (ks < C / 16) {
        const float* src = a.proj + (row0 + i0 + k) * NPROJ + OFF_Q + h * C + kk;
        const float2 x0 = __ldg(reinterpret_cast<const float2*>(src)), x1 = __ldg(reinterpret_cast<const float2*>(src + 8));
        v[half * 2] = x0.x * V2_QSCALE; v[half * 2 + 1] = x0.y * V2_QSCALE;
        v[4 + half * 2] = x1.x * V2_QSCALE; v[4 + half * 2 + 1] = x1.y * V2_QSCALE;
      } else {
        const float* src = s_qp + (k * H + h) * V2_KPW;
        const int e = kk - C;
        v[half * 2] = e < V2_KPW ? src[e] * ch : 0.f;
        v[half * 2 + 1] = e + 1 < V2_KPW ? src[e + 1] * ch : 0.f;
        v[4 + half * 2] = e + 8 < V2_KPW ? src[e + 8] * ch : 0.f;
        v[4 + half * 2 + 1] = e + 9 < V2_KPW ? src[e + 9] * ch : 0.f;
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

// ---- persistent, double-buffered blob packer: the K' / V' fragments, key bias, key mask and key translations, with the
// rows of the NEXT key tile landing in shared memory (bulk copies on an mbarrier) while the current tile is being cut
// into fragments, one CTA per SM walking the tiles with stride gridDim.x.  The Q' tiles keep the kernel above.
constexpr int P4_THREADS = 768;                                    // one CTA per SM: 24 warps cut fragments
constexpr int P4_STAGE = V2_TK * P3_KVP + V2_TK * 480;            // floats per stage: k|v rows, local kv points
constexpr int P4_OFF_KP = 2 * P4_STAGE;
constexpr int P4_OFF_VP = P4_OFF_KP + V2_TK * H * V2_KPW;        // s_kp rows: [B | t | b_p] (30 floats)
constexpr int P4_OFF_FR = P4_OFF_VP + V2_TK * H * PV * 3;
constexpr int P4_OFF_BAR = ((P4_OFF_FR + V2_TK * 12 + 1) & ~1);     // 8-byte aligned mbarriers
constexpr int P4_SMEM = (P4_OFF_BAR + 4) * 4;
static_assert((P3_KVP * 4) % 16 == 0 && (P4_STAGE * 4) % 16 == 0 && (V2_TK * P3_KVP * 4) % 16 == 0, "bulk-copy alignment");
static_assert(P4_SMEM <= 232448, "shared memory budget");

__global__ void __launch_bounds__(P4_THREADS, 1) ipa_pack4_kernel(IpaPack3Args a) {
  extern __shared__ __align__(16) float p4[];
  const int L = a.L, tid = threadIdx.x;
  const int nblob = a.B * a.JT;
  float* s_kp = p4 + P4_OFF_KP;
  float* s_vp = p4 + P4_OFF_VP;
  float* s_fr = p4 + P4_OFF_FR;
  const uint32_t bar0 = smem_u32(p4 + P4_OFF_BAR);
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int blob, int stage) {                    // one thread: rows of key tile `blob` -> stage
    const int b = blob / a.JT, jt = blob - b * a.JT, j0 = jt * V2_TK;
    const int nk = min(V2_TK, L - j0);
    const float* src = a.proj + ((size_t)b * L + j0) * NPROJ;
    const uint32_t dkv = smem_u32(p4 + stage * P4_STAGE), dloc = dkv + V2_TK * P3_KVP * 4;
    const uint32_t bar = bar0 + 8 * stage;
    mbar_arrive_expect_tx(bar, nk * (2 * H * C * 4 + 480 * 4));
    for (int k = 0; k < nk; ++k) {
      bulk_g2s(dkv + k * P3_KVP * 4, src + (size_t)k * NPROJ + OFF_KV, 2 * H * C * 4, bar);
      bulk_g2s(dloc + k * 480 * 4, src + (size_t)k * NPROJ + OFF_KVP, 480 * 4, bar);
    }
  };
  if (tid == 0 && (int)blockIdx.x < nblob) issue(blockIdx.x, 0);
  int n = 0;
  for (int blob = blockIdx.x; blob < nblob; blob += gridDim.x, ++n) {
    const int stage = n & 1;
    const int b = blob / a.JT, jt = blob - b * a.JT, j0 = jt * V2_TK;
    const size_t row0 = (size_t)b * L;
    const int nk = min(V2_TK, L - j0);
    // every thread has left the previous tile (its reads of stage ^ 1 included): refill that stage
    if (tid == 0 && blob + (int)gridDim.x < nblob) issue(blob + gridDim.x, stage ^ 1);
    if (tid < V2_TK * 12) {
      const int k = tid / 12, e = tid - k * 12;
      s_fr[tid] = k < nk ? (e < 9 ? a.rot[(row0 + j0 + k) * 9 + e] : a.trans[(row0 + j0 + k) * 3 + e - 9]) : 0.f;
    }
    mbar_wait_cta(bar0 + 8 * stage, (n >> 1) & 1);
    __syncthreads();
    const float* s_kv = p4 + stage * P4_STAGE;
    const float* s_loc = s_kv + V2_TK * P3_KVP;
    for (int i = tid; i < V2_TK * H * P3_NP; i += P4_THREADS) {        // key points: rotate; value points: rotate + translate
      const int k = i / (H * P3_NP), r = i - k * (H * P3_NP), h = r / P3_NP, pnt = r - h * P3_NP;
      const bool on = k < nk;                                          // keys past the end stay exactly zero
      const float* lp = s_loc + k * 480 + h * P3_NP + pnt;            // [xyz][h][20]
      const float lx = on ? lp[0] : 0.f, ly = on ? lp[H * P3_NP] : 0.f, lz = on ? lp[2 * H * P3_NP] : 0.f;
      const float* R = s_fr + k * 12;
      const bool isk = pnt < PQ;
      float* dst = isk ? s_kp + (k * H + h) * V2_KPW + 6 + pnt * 3 : s_vp + (k * H + h) * (PV * 3) + (pnt - PQ) * 3;
      const float tx = isk ? 0.f : R[9], ty = isk ? 0.f : R[10], tz = isk ? 0.f : R[11];
      dst[0] = on ? R[0] * lx + R[1] * ly + R[2] * lz + tx : 0.f;
      dst[1] = on ? R[3] * lx + R[4] * ly + R[5] * lz + ty : 0.f;
      dst[2] = on ? R[6] * lx + R[7] * ly + R[8] * lz + tz : 0.f;
    }
    __syncthreads();
    if (tid < V2_TK * H * 3) {                                         // B_j = sum_p b_jp and t_j in front of the points
      const int k = tid / (H * 3), r = tid - k * (H * 3), h = r / 3, x = r - h * 3;
      float* row = s_kp + (k * H + h) * V2_KPW;
      float sum = 0.f;
#pragma unroll
      for (int pnt = 0; pnt < PQ; ++pnt) sum += row[6 + pnt * 3 + x];
      row[x] = sum;
      row[3 + x] = s_fr[k * 12 + 9 + x];                               // 0 for keys past the end
    }
    __syncthreads();
    unsigned char* out = a.blobs + (size_t)blob * V2_BLOB;
#pragma unroll 5
    for (int idx = tid; idx < H * V2_KS * 32; idx += P4_THREADS) {     // K' [h][ks][lane]
      const int lane = idx & 31, r = idx >> 5, ks = r % V2_KS, h = r / V2_KS;
      const int g = lane >> 2, t = lane & 3, kk = ks * 16 + 2 * t;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (g < nk) {                                                    // rows past the end were not copied
        if (ks < C / 16) {
          const float* src = s_kv + g * P3_KVP + h * 2 * C + kk;
          const float2 x0 = *reinterpret_cast<const float2*>(src), x1 = *reinterpret_cast<const float2*>(src + 8);
          v[0] = x0.x; v[1] = x0.y; v[2] = x1.x; v[3] = x1.y;
        } else {
          const float* src = s_kp + (g * H + h) * V2_KPW;
          const int e = kk - C;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ee = e + (q & 1) + (q >> 1) * 8;
            v[q] = ee < V2_KPW ? src[ee] : 0.f;
          }
        }
      }
      uint4 o;
      split_pair(v[0], v[1], o.x, o.z);
      split_pair(v[2], v[3], o.y, o.w);
      reinterpret_cast<uint4*>(out)[idx] = o;
    }
#pragma unroll 3
    for (int idx = tid; idx < H * V2_VNT * 32; idx += P4_THREADS) {    // V' [h][nt][lane]
      const int lane = idx & 31, r = idx >> 5, nt = r % V2_VNT, h = r / V2_VNT;
      const int g = lane >> 2, t = lane & 3, nn = nt * 8 + g, k0 = 2 * t;
      float v0 = 0.f, v1 = 0.f;
      if (nn < C) {
        if (k0 < nk) v0 = s_kv[k0 * P3_KVP + h * 2 * C + C + nn];
        if (k0 + 1 < nk) v1 = s_kv[(k0 + 1) * P3_KVP + h * 2 * C + C + nn];
      } else if (nn < C + PV * 3) {
        v0 = s_vp[(k0 * H + h) * (PV * 3) + nn - C]; v1 = s_vp[((k0 + 1) * H + h) * (PV * 3) + nn - C];
      }
      uint2 o;
      split_pair(v0, v1, o.x, o.y);
      reinterpret_cast<uint2*>(out + V2_OFF_V)[idx] = o;
    }
    if (tid < (H + 1) * V2_TK) {                                       // key bias [h][key], key mask [key]
      const int key = tid % V2_TK, h = tid / V2_TK, j = j0 + key;
      float* dst = reinterpret_cast<float*>(out + V2_OFF_KB);
      if (h == H) {
        dst[H * V2_TK + key] = (j < L) ? a.mask[row0 + j] : 0.f;
      } else {
        const float* kp = s_kp + (key * H + h) * V2_KPW;               // [B | t | b_p]
        float sq = 0.f;
#pragma unroll
        for (int e = 0; e < PQ * 3; ++e) sq = fmaf(kp[6 + e], kp[6 + e], sq);
        const float tb = kp[0] * kp[3] + kp[1] * kp[4] + kp[2] * kp[5];
        dst[h * V2_TK + key] = -a.head_w[h] * (tb + 0.5f * sq);        // -c_h t_j.B_j - 1/2 c_h sum_p |b_jp|^2
      }
    } else if (tid >= 128 && tid < 128 + V2_TK * 4) {                  // key translations [key][x y z 0]
      const int key = (tid - 128) >> 2, x = (tid - 128) & 3;
      reinterpret_cast<float*>(out + V2_OFF_T)[key * 4 + x] = x < 3 ? s_fr[key * 12 + 9 + x] : 0.f;
    }
    __syncthreads();                                                   // s_kp / s_vp / s_fr / this stage are free again
  }
}

// D(16x8, fp32) += A(16x8, fp16, row) * B(8x8, fp16, col)
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
constexpr int V2_BP = 10;                                    // pitch of the bias rows [h][i][10]

struct Ipa2Args {
  IpaArgs a;
  const unsigned char* blobs; const uint4* Qp;
  int JT, IT;
};

size_t ipa_workspace_bytes(int B, int L) { return ipa_v2_workspace_bytes(B, L); }

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
// shared-memory layout; DEC = decoupled pair-warp pipeline (variant 4): V4_SLOTS row slots of 2 KB per pair warp
// (one mbarrier each) instead of 2 stages of 2 rows, paid for by moving the Q' lo fragments to tensor memory
constexpr int V4_SLOTS = 6;
#ifndef V4_HEAD_REGS
#define V4_HEAD_REGS 168
#endif
// Every head warp owns the K' / V' slice of its head: it waits on its own mbarrier and refills the slice itself as
// soon as IT is done with it - no CTA-wide free / full hand-off, the eight warps drift freely.
constexpr int V3_KH = V2_KS * 32 * 16 + V2_TK * 4 + V2_TK * 4 + V2_BLOB_T;   // 5312: K' fragments | key bias [8] | key mask [8] | key translations [8][4]
constexpr int V3_VH = V2_VNT * 32 * 8;                            // 5376: V' fragments
constexpr int V3_PH = 68;            // uint2 per head in a P tile (64 used; pitch = 8 words mod 32)
template <bool DEC>
struct V3L {
  static constexpr int ZWARP = DEC ? V4_SLOTS * 2048 : 2 * V3_ZSTAGE_W;  // z bytes owned by one pair warp
  static constexpr int NZBAR = DEC ? V4_SLOTS : 2;                 // z mbarriers per pair warp
  static constexpr int SM_BLOB = 0;                                // [8 h][V3_KH] K' slices, then [8 h][V3_VH] V' slices
  static constexpr int SM_VB = SM_BLOB + H * V3_KH;
  static constexpr int SM_Z = (SM_VB + H * V3_VH + 1023) & ~1023;  // [8 warps][ZWARP]
  static constexpr int SM_QLO = SM_Z + 8 * ZWARP;                  // [8 h][10 ks][32 lanes] uint4 (not DEC)
  static constexpr int SM_WB = SM_QLO + (DEC ? 0 : H * V2_KS * 32 * 16);   // [4 ks][32 lanes] uint4 {hi b0, hi b1, lo b0, lo b1}
  static constexpr int SM_BIAS = SM_WB + 4 * 32 * 16;              // [2][8 h][16 i][10]
  static constexpr int SM_P = SM_BIAS + 2 * H * V2_TQ * V2_BP * 4; // [2][8 h][V3_PH] uint2 {P hi, P lo}: [16 i][4 key pairs] + pad
  static constexpr int SM_ALPHA = SM_P + 2 * H * V3_PH * 8;        // [2][8 h][16 i]
  static constexpr int SM_L = SM_ALPHA + 2 * H * V2_TQ * 4;        // [8 h][16 i]
  static constexpr int SM_BAR = SM_L + H * V2_TQ * 4;              // K' full [8 h], V' full [8 h], [8 warps][NZBAR] z mbarriers
  static constexpr int SM_HBAR = SM_BAR + 128 + 8 * NZBAR * 8;     // bias-ready [2] and P-ready [2] mbarriers
  static constexpr int SM_TMEM = SM_HBAR + 32;                     // tensor-memory base address (DEC)
  static constexpr int SMEM = SM_TMEM + 16;
  // epilogue staging of W_dz [16 d][68]: the Q' lo region, or (DEC) the z region behind the o_pt scratch
  static constexpr int SM_WDZ = DEC ? SM_Z + H * V2_TQ * 40 * 4 : SM_QLO;
  static_assert(!DEC || H * V2_TQ * 40 * 4 + 16 * V2_ZP * 4 <= 8 * ZWARP, "W_dz staging fits behind the o_pt scratch");
  static_assert(SM_Z % 1024 == 0 && SM_QLO % 16 == 0 && SM_P % 16 == 0 && SM_BAR % 8 == 0, "alignment");
  static_assert(SMEM <= 232448, "shared memory budget");
  // epilogue re-use of the loop buffers
  static_assert(V2_TQ * H * V2_ZP * 4 <= H * (V3_KH + V3_VH), "o_pair_raw fits in the blob region");
  static_assert(H * V2_TQ * 40 * 4 <= 8 * ZWARP, "o_pt scratch fits in the z region");
};

// four 8x8 b16 matrices, transposed on the way in; lane l supplies the address of row l & 7 of matrix l >> 3
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
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

template <bool DEC>
__global__ void __launch_bounds__(V3_THREADS, 1) ipa_attention_v3_kernel(const __grid_constant__ Ipa3Args args) {
  using LT = V3L<DEC>;
  const Ipa2Args& p = args.p;
  extern __shared__ __align__(128) unsigned char smem[];
  const unsigned char* blob = smem + LT::SM_BLOB;
  float* sbias = reinterpret_cast<float*>(smem + LT::SM_BIAS);
  uint2* sP = reinterpret_cast<uint2*>(smem + LT::SM_P);
  float* salpha = reinterpret_cast<float*>(smem + LT::SM_ALPHA);
  float* sl = reinterpret_cast<float*>(smem + LT::SM_L);
  const uint32_t bar_kfull = smem_u32(smem + LT::SM_BAR), bar_vfull = bar_kfull + 64;   // [8 h] each
  // hand-offs between the warp groups: per-tile-parity mbarriers instead of CTA-wide named barriers, so the eight head
  // warps are not forced into lockstep (one warp's softmax runs under another's MMAs)
  const uint32_t bar_bias = smem_u32(smem + LT::SM_HBAR), bar_p = bar_bias + 16;
  const IpaArgs& a = p.a;
  const int L = a.L, JT = p.JT;
  const int b = blockIdx.y, it = blockIdx.x, i0 = it * V2_TQ;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const size_t rowb = (size_t)b * L;
  const float sc_b = 0.5773502691896257f;              // sqrt(1/3)
  const unsigned char* gblob = p.blobs + (size_t)b * JT * V2_BLOB;

  if (tid == 0) {
    for (int q = 0; q < 16; ++q) mbar_init(bar_kfull + 8 * q, 1);
    for (int q = 0; q < 4; ++q) mbar_init(bar_bias + 8 * q, 8);   // one arrival per producing warp
    for (int q = 0; q < 8 * LT::NZBAR; ++q) mbar_init(bar_kfull + 128 + 8 * q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    tma_prefetch_desc(&args.tm_z);
  }
  // W_b as B fragments of the pair-bias MMA (rows n = head, k = channel; sqrt(1/3) folded in)
  if (tid < 128) {
    const int ks = tid >> 5, gg = (tid & 31) >> 2, tt = tid & 3;
    // K slots (2t, 2t+1 | 2t+8, 2t+9) of K step ks carry channels 16 ks + 4t .. 4t+3, so that the matching A
    // fragment is ONE 16-byte load per pair
    const float* w = a.w_b + gg * CZ + ks * 16 + 4 * tt;
    uint4 o;
    split_pair(w[0] * sc_b, w[1] * sc_b, o.x, o.z);
    split_pair(w[2] * sc_b, w[3] * sc_b, o.y, o.w);
    reinterpret_cast<uint4*>(smem + LT::SM_WB)[tid] = o;
  }
  // lo halves of the Q' fragments (the hi halves stay in the head warps' registers): shared memory, or (DEC) tensor
  // memory - 40 columns per head warp, written below by the warp that reads them back
  if constexpr (!DEC) {
    for (int idx = tid; idx < H * V2_KS * 32; idx += V3_THREADS) {
      const int hh = idx / (V2_KS * 32), rem = idx % (V2_KS * 32);
      reinterpret_cast<uint4*>(smem + LT::SM_QLO)[idx] =
          p.Qp[(((size_t)b * H + hh) * p.IT + it) * V2_QTILE_U4 + rem * 2 + 1];
    }
  } else {
    if (warp == 0) tmem_alloc_cta(smem_u32(smem + LT::SM_TMEM), 256);
    tc_fence_before();
  }
  __syncthreads();
  uint32_t tmem_base = 0;
  if constexpr (DEC) {
    tc_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + LT::SM_TMEM);
  }

  if (warp < 8) {
    // ========================================== head warps ======================================
    if constexpr (DEC) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(V4_HEAD_REGS));
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 176;\n");
    const int h = warp;
    uint4 qh[DEC ? 1 : V2_KS];                        // Q' hi fragments: registers, or (DEC) tensor memory
    const uint4* qlo = reinterpret_cast<const uint4*>(smem + LT::SM_QLO) + (h * V2_KS) * 32 + lane;
    // (DEC) tensor-memory address of this warp's Q' fragments: its own lane quadrant, 80 columns per warp
    // ([ks]{hi x4, lo x4}), which leaves the head warps 40 registers lighter and the pair warps 104 registers
    const uint32_t tq = tmem_base + ((32u * (warp & 3)) << 16) + 80u * (warp >> 2);
    {
      const uint4* qp = p.Qp + (((size_t)b * H + h) * p.IT + it) * V2_QTILE_U4;
      if constexpr (!DEC) {
#pragma unroll
        for (int ks = 0; ks < V2_KS; ++ks) qh[ks] = qp[(ks * 32 + lane) * 2];
      } else {
#pragma unroll
        for (int ks = 0; ks < V2_KS; ++ks) {
          const uint4 hi = qp[(ks * 32 + lane) * 2], lo = qp[(ks * 32 + lane) * 2 + 1];
          const uint32_t r[8] = {hi.x, hi.y, hi.z, hi.w, lo.x, lo.y, lo.z, lo.w};
          tmem_st8(tq + 8 * ks, r);
        }
        tc_wait_st();
      }
    }
    const float bbias = sc_b * a.b_b[h];
    const float mi_lo = (i0 + g < L) ? a.mask[rowb + i0 + g] : 0.f;
    const float mi_hi = (i0 + g + 8 < L) ? a.mask[rowb + i0 + g + 8] : 0.f;
    // frame translations of this lane's two query rows and -P/2 c_h: the |t_i - t_j|^2 part of the point term is
    // evaluated here in plain fp32, per pair
    float ti_lo[3], ti_hi[3];
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      ti_lo[x] = (i0 + g < L) ? a.trans[(rowb + i0 + g) * 3 + x] : 0.f;
      ti_hi[x] = (i0 + g + 8 < L) ? a.trans[(rowb + i0 + g + 8) * 3 + x] : 0.f;
    }
    const float c_dist = -0.5f * PQ * a.head_w[h];
    const unsigned char* kmine = smem + LT::SM_BLOB + h * V3_KH;
    const unsigned char* vmine = smem + LT::SM_VB + h * V3_VH;
    const float* skb = reinterpret_cast<const float*>(kmine + V2_KS * 32 * 16);   // key bias [8] then key mask [8]
    const uint32_t my_kfull = bar_kfull + 8 * h, my_vfull = bar_vfull + 8 * h;
    auto issue_k = [&](int jt) {                       // one lane: this head's K' fragments, key bias, key mask
      const unsigned char* src = gblob + (size_t)jt * V2_BLOB;
      const uint32_t dst = smem_u32(kmine);
      mbar_arrive_expect_tx(my_kfull, V3_KH);
      bulk_g2s(dst, src + h * (V2_KS * 32 * 16), V2_KS * 32 * 16, my_kfull);
      bulk_g2s(dst + V2_KS * 32 * 16, src + V2_OFF_KB + h * (V2_TK * 4), V2_TK * 4, my_kfull);
      bulk_g2s(dst + V2_KS * 32 * 16 + V2_TK * 4, src + V2_OFF_M, V2_TK * 4 + V2_BLOB_T, my_kfull);   // mask | translations
    };
    auto issue_v = [&](int jt) {
      mbar_arrive_expect_tx(my_vfull, V3_VH);
      bulk_g2s(smem_u32(vmine), gblob + (size_t)jt * V2_BLOB + V2_OFF_V + h * V3_VH, V3_VH, my_vfull);
    };
    float O[V2_VNT][4];
#pragma unroll
    for (int n = 0; n < V2_VNT; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
    if (lane == 0) { issue_k(0); issue_v(0); }
    for (int jt = 0; jt < JT; ++jt) {
      const int j0 = jt * V2_TK, par = jt & 1;
      // ---- S = Q' K'^T (16 rows x 8 keys)
      mbar_wait_cta(my_kfull, par);
      float S[4];
      float kbv[4], mjv[2];                            // kbv: key bias + bias of linear_b + distance term, rows g | g+8
      {
        float Sc[4][4];                                // four independent accumulation chains (7-8 MMAs each)
#pragma unroll
        for (int q = 0; q < 4; ++q) { Sc[q][0] = Sc[q][1] = Sc[q][2] = Sc[q][3] = 0.f; }
        const uint4* kp = reinterpret_cast<const uint4*>(kmine) + lane;
        if constexpr (!DEC) {
#pragma unroll
          for (int ks = 0; ks < V2_KS; ++ks) {
            const uint4 k0 = kp[ks * 32];
            const uint4 q_lo = qlo[ks * 32];
            const uint32_t ah[4] = {qh[ks].x, qh[ks].y, qh[ks].z, qh[ks].w};
            const uint32_t al[4] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w};
            mma16816(Sc[(3 * ks) & 3], al, k0.x, k0.y);
            mma16816(Sc[(3 * ks + 1) & 3], ah, k0.x, k0.y);
            mma16816(Sc[(3 * ks + 2) & 3], ah, k0.z, k0.w);
          }
        } else {
          // Q' comes back from tensor memory one K step at a time; the load of the next step is in flight under the
          // MMAs of this one
          uint32_t qq[2][8];
          tmem_ld8(tq, qq[0]);
          tc_wait_ld();
#pragma unroll
          for (int ks = 0; ks < V2_KS; ++ks) {
            if (ks + 1 < V2_KS) tmem_ld8(tq + 8 * (ks + 1), qq[(ks + 1) & 1]);
            const uint4 k0 = kp[ks * 32];
            const uint32_t ah[4] = {qq[ks & 1][0], qq[ks & 1][1], qq[ks & 1][2], qq[ks & 1][3]};
            const uint32_t al[4] = {qq[ks & 1][4], qq[ks & 1][5], qq[ks & 1][6], qq[ks & 1][7]};
            mma16816(Sc[(3 * ks) & 3], al, k0.x, k0.y);
            mma16816(Sc[(3 * ks + 1) & 3], ah, k0.x, k0.y);
            mma16816(Sc[(3 * ks + 2) & 3], ah, k0.z, k0.w);
            tc_wait_ld();
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) S[e] = (Sc[0][e] + Sc[1][e]) + (Sc[2][e] + Sc[3][e]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float4 tj = *reinterpret_cast<const float4*>(skb + 2 * V2_TK + 4 * (2 * t + e));
          const float ax = ti_lo[0] - tj.x, ay = ti_lo[1] - tj.y, az = ti_lo[2] - tj.z;
          const float bx = ti_hi[0] - tj.x, by = ti_hi[1] - tj.y, bz = ti_hi[2] - tj.z;
          const float kb = skb[2 * t + e] + bbias;
          kbv[e] = kb + c_dist * (ax * ax + ay * ay + az * az);
          kbv[2 + e] = kb + c_dist * (bx * bx + by * by + bz * bz);
          mjv[e] = skb[V2_TK + 2 * t + e];
        }
      }
      __syncwarp();                                    // this warp is done with its K' / key bias / mask of the tile
      if (lane == 0 && jt + 1 < JT) issue_k(jt + 1);
      // ---- logits and online softmax (row g: S[0..1], row g+8: S[2..3]; keys 2t, 2t+1)
      mbar_wait_cta(bar_bias + 8 * par, (jt >> 1) & 1);    // pair bias of this tile is in sbias[par]
      uint32_t ph0, pl0, ph1, pl1;
      {
        const float* bs = sbias + par * (H * V2_TQ * V2_BP);
        float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = 2 * t + e;
          const bool valid = (j0 + j < L);
          const float b_lo = bs[(h * V2_TQ + g) * V2_BP + j], b_hi = bs[(h * V2_TQ + g + 8) * V2_BP + j];
          float x_lo = S[e] + b_lo + kbv[e] + 1e5f * (mi_lo * mjv[e] - 1.f);
          float x_hi = S[2 + e] + b_hi + kbv[2 + e] + 1e5f * (mi_hi * mjv[e] - 1.f);
          x_lo = valid ? x_lo : -INFINITY;
          x_hi = valid ? x_hi : -INFINITY;
          S[e] = x_lo; S[2 + e] = x_hi;
          mx_lo = fmaxf(mx_lo, x_lo); mx_hi = fmaxf(mx_hi, x_hi);
        }
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
        mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
        mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
        // Lazy rescale: the running maximum only moves when some row of the warp exceeds it by more than 8
        // (P <= e^8 stays far inside fp16 / fp32 range); otherwise alpha = 1 and the 84 accumulator multiplies of
        // this warp - and the pair warps' - are skipped.  The final O / l is unchanged mathematically.
#ifndef V3_NO_LAZY
        const bool resc = __any_sync(0xffffffffu, (mx_lo > m_lo + 8.f) || (mx_hi > m_hi + 8.f));
#else
        const bool resc = true;
#endif
        float al_lo = 1.f, al_hi = 1.f;
        if (resc) {
          const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
          al_lo = __expf(m_lo - mn_lo); al_hi = __expf(m_hi - mn_hi);         // exp(-inf) = 0 on the first tile
          m_lo = mn_lo; m_hi = mn_hi;
        }
        float ps_lo = 0.f, ps_hi = 0.f;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          // ex2.approx(x log2 e): relative error ~|x| 2^-24 <= 3e-6 over the range of a softmax row, an order of
          // magnitude under the split-precision error of the logits themselves
          const float p_lo = __expf(S[e] - m_lo), p_hi = __expf(S[2 + e] - m_hi);
          S[e] = p_lo; S[2 + e] = p_hi;
          ps_lo += p_lo; ps_hi += p_hi;
        }
        split_pair(S[0], S[1], ph0, pl0);              // a0: row g,   keys 2t, 2t+1
        split_pair(S[2], S[3], ph1, pl1);              // a1: row g+8
        {
          uint2* Pw = sP + (par * H + h) * V3_PH;      // [row][key pair] {hi, lo}: what the pair warps' B fragments hold
          Pw[g * 4 + t] = make_uint2(ph0, pl0);
          Pw[(g + 8) * 4 + t] = make_uint2(ph1, pl1);
        }
        l_lo = l_lo * al_lo + ps_lo;
        l_hi = l_hi * al_hi + ps_hi;
        if (t == 0) {
          salpha[par * (H * V2_TQ) + h * V2_TQ + g] = al_lo;
          salpha[par * (H * V2_TQ) + h * V2_TQ + g + 8] = al_hi;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + 8 * par);   // P / alpha of this tile are in sP[par], salpha[par]
        if (resc) {
#pragma unroll
          for (int n = 0; n < V2_VNT; ++n) { O[n][0] *= al_lo; O[n][1] *= al_lo; O[n][2] *= al_hi; O[n][3] *= al_hi; }
        }
      }
      // ---- O += P [V | v_pts]   (m16n8k8: K = the 8 keys of the tile)
      mbar_wait_cta(my_vfull, par);
      {
        // HMMA m16n8k8 occupies the tensor pipe as long as m16n8k16 (8 cycles per sub-core, measured with
        // scripts/ubench/hmma_rate.cu), so two of the three split-precision products share one K = 16 MMA
        const uint32_t pa[4] = {ph0, ph1, pl0, pl1};
        const uint2* vp = reinterpret_cast<const uint2*>(vmine) + lane;
#pragma unroll
        for (int n = 0; n < V2_VNT; ++n) {
          const uint2 v = vp[n * 32];
          mma16816(O[n], pa, v.x, v.x);                // K = [P hi | P lo] x [V hi ; V hi]
          mma1688(O[n], ph0, ph1, v.y);                // P hi x V lo
        }
      }
      __syncwarp();
      if (lane == 0 && jt + 1 < JT) issue_v(jt + 1);
    }
    // ---- head epilogue: normalise, write o, stage o_pt
    l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
    l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
    const float il_lo = 1.0f / l_lo, il_hi = 1.0f / l_hi;
    if (t == 0) { sl[h * V2_TQ + g] = il_lo; sl[h * V2_TQ + g + 8] = il_hi; }
    named_sync(5, V3_THREADS);                         // every warp has left the tile loop: z region and blob are free
    float* spt = reinterpret_cast<float*>(smem + LT::SM_Z);   // [8 h][16 i][40] normalised global-frame o_pt
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
    if constexpr (DEC) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(256 - V4_HEAD_REGS));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 80;\n");
    const int pw = warp - 8, r0 = 2 * pw;             // this warp owns query rows r0, r0 + 1
    const uint32_t zbase = smem_u32(smem + LT::SM_Z + pw * LT::ZWARP);
    unsigned char* zgen = smem + LT::SM_Z + pw * LT::ZWARP;
    const uint32_t zbar = bar_kfull + 128 + 8 * LT::NZBAR * pw;
    const uint4* wbf = reinterpret_cast<const uint4*>(smem + LT::SM_WB) + lane;
    float acc[2][4][4];   // o_pair_raw^T: [row][m-tile of 16 channels][C fragment: (ch g | g+8) x (heads 2t, 2t+1)]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) { acc[r][mt][0] = acc[r][mt][1] = acc[r][mt][2] = acc[r][mt][3] = 0.f; }
    // ---- pair bias for rows r0 (z tile at za), r0+1 (zb) x 8 keys, all heads -> sbias[par].  The fp16 hi / lo split
    //      of z that feeds the MMA is also written back IN PLACE: once both K steps of a 32-channel half have been
    //      read, that 1 KB half of each row ([8 keys][32 ch] fp32) is dead and receives [hi | lo][8 keys][32 ch] fp16
    //      (16-byte chunks XOR-swizzled by key >> 1), from which the o_pair phase fetches its transposed A
    //      fragments with ldmatrix.trans instead of re-reading and re-splitting the fp32 tile.
    //      MMA row g carries key kq = (g >> 1) + 4 (g & 1): the two keys of a quarter-warp then sit in different
    //      halves of the 128-byte swizzle and every 16-byte load / 8-byte store below is conflict-free.
    const int kq = (g >> 1) + 4 * (g & 1);
    const int st_off = kq * 64 + 8 * (t & 1);          // this lane's 8 bytes inside a [8 keys][64 B] fp16 block
    auto bias_phase = [&](unsigned char* za, unsigned char* zb, int par) {
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hh[2][4], ll[2][4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ks = 2 * hf + e;
          const int zo = v3_zoff(0, kq, ks * 16 + 4 * t);
          const float4 xa = *reinterpret_cast<const float4*>(za + zo);        // pair (row r0,   key kq), 4 channels
          const float4 xb = *reinterpret_cast<const float4*>(zb + zo);        // pair (row r0+1, key kq)
          split_pair(xa.x, xa.y, hh[e][0], ll[e][0]);
          split_pair(xb.x, xb.y, hh[e][1], ll[e][1]);
          split_pair(xa.z, xa.w, hh[e][2], ll[e][2]);
          split_pair(xb.z, xb.w, hh[e][3], ll[e][3]);
          const uint4 w = wbf[ks * 32];
          mma16816(c0, ll[e], w.x, w.y);
          mma16816(c1, hh[e], w.x, w.y);
          mma16816(c0, hh[e], w.z, w.w);
        }
        __syncwarp();                                  // every lane has read this half of both rows
#pragma unroll
        for (int e = 0; e < 2; ++e) {                  // channels 16 e + 4t .. 4t+3 of the half: 16-byte chunk 2e + (t >> 1)
          const int sw = (((2 * e + (t >> 1)) ^ (kq >> 1)) & 3) * 16 + st_off + hf * 1024;
          *reinterpret_cast<uint2*>(za + sw) = make_uint2(hh[e][0], hh[e][2]);
          *reinterpret_cast<uint2*>(za + sw + 512) = make_uint2(ll[e][0], ll[e][2]);
          *reinterpret_cast<uint2*>(zb + sw) = make_uint2(hh[e][1], hh[e][3]);
          *reinterpret_cast<uint2*>(zb + sw + 512) = make_uint2(ll[e][1], ll[e][3]);
        }
      }
      float* bs = sbias + par * (H * V2_TQ * V2_BP);
      bs[((2 * t) * V2_TQ + r0) * V2_BP + kq] = c0[0] + c1[0];
      bs[((2 * t + 1) * V2_TQ + r0) * V2_BP + kq] = c0[1] + c1[1];
      bs[((2 * t) * V2_TQ + r0 + 1) * V2_BP + kq] = c0[2] + c1[2];
      bs[((2 * t + 1) * V2_TQ + r0 + 1) * V2_BP + kq] = c0[3] + c1[3];
      __syncwarp();                                    // the fp16 copy is complete before any lane's ldmatrix
    };
    // ---- o_pair_raw^T[c, h] = alpha_h * old + sum_j z[j, c] P[h, j] for query row r0 + r (fp16 copy of its z tile at
    //      zr): M = 16 channels, N = 8 heads, K = [z hi | z lo] x [P hi ; P hi] (m16n8k16) + z hi x P lo (m16n8k8).
    //      ldmatrix.x4.trans: lanes 8m..8m+7 address the key rows of matrix m = {hi ch 0-7, hi ch 8-15, lo ch 0-7,
    //      lo ch 8-15} of a 16-channel group; transposed, thread (g, t) receives (channel g; keys 2t, 2t+1).
    const int lm_k = lane & 7, lm_m = lane >> 3;
    const int lm_base = (lm_m >> 1) * 512 + lm_k * 64;
    const int lm_even = lm_base + ((((lm_m & 1)) ^ (lm_k >> 1)) & 3) * 16;       // channel groups 0, 2: chunks 0, 1
    const int lm_odd = lm_base + (((2 + (lm_m & 1)) ^ (lm_k >> 1)) & 3) * 16;    // channel groups 1, 3: chunks 2, 3
    auto opair_row = [&](const unsigned char* zr, int r, int par) {
      const uint2 pb = sP[(par * H + g) * V3_PH + (r0 + r) * 4 + t];   // B = P^T: (keys 2t, 2t+1; head g) {hi, lo}
      const float* al = salpha + par * (H * V2_TQ);
      const float a0 = al[(2 * t) * V2_TQ + r0 + r], a1 = al[(2 * t + 1) * V2_TQ + r0 + r];
#ifndef V3_NO_LAZY
      const bool resc = __any_sync(0xffffffffu, (a0 != 1.f) || (a1 != 1.f));
#else
      const bool resc = true;
#endif
      const uint32_t zs = smem_u32(zr);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        uint32_t za[4];                                // {hi ch g, hi ch g+8, lo ch g, lo ch g+8} of channels 16 mt ..
        ldmatrix_x4_trans(za, zs + (mt >> 1) * 1024 + ((mt & 1) ? lm_odd : lm_even));
        float (&d)[4] = r == 0 ? acc[0][mt] : acc[1][mt];
        if (resc) { d[0] *= a0; d[1] *= a1; d[2] *= a0; d[3] *= a1; }
        mma16816(d, za, pb.x, pb.x);
        mma1688(d, za[0], za[1], pb.y);
      }
    };
    if constexpr (!DEC) {
      auto issue_z = [&](int jt) {                       // one lane: 4 boxes = (2 rows) x (2 channel halves)
        const int st = jt & 1;
        mbar_arrive_expect_tx(zbar + 8 * st, V3_ZSTAGE_W);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          tma_load_3d(zbase + st * V3_ZSTAGE_W + q * 1024, &args.tm_z, 32 * (q & 1), jt * V2_TK,
                      (int)rowb + i0 + r0 + (q >> 1), zbar + 8 * st);
      };
      if (lane == 0) issue_z(0);
      for (int jt = 0; jt < JT; ++jt) {
        const int par = jt & 1;
        __syncwarp();                                    // every lane is done with the other stage (tile jt - 1)
        if (lane == 0 && jt + 1 < JT) issue_z(jt + 1);
        mbar_wait_cta(zbar + 8 * par, (jt >> 1) & 1);        // z tile jt (this warp's rows) has landed
        unsigned char* zt = zgen + par * V3_ZSTAGE_W;
        bias_phase(zt, zt + 2048, par);
        if (lane == 0) mbar_arrive(bar_bias + 8 * par);
        mbar_wait_cta(bar_p + 8 * par, (jt >> 1) & 1);       // the head warps have published P / alpha of the tile
        opair_row(zt, 0, par);
        opair_row(zt + 2048, 1, par);
      }
    } else {
      // Decoupled pipeline: the bias of tile jt + 1 is computed BEFORE the o_pair accumulation of tile jt, so the
      // head warps find their bias waiting and the loop-carried chain bias -> softmax -> P -> o_pair -> next bias
      // is broken.  Row tiles (item n = 2 jt + r, [8 keys][64 ch], 2 KB) cycle through V4_SLOTS = 6 slots: two tiles
      // resident, one in flight with a full iteration of lead; the slot of an item is refilled with item
      // n + V4_SLOTS as soon as its o_pair part is done.
      const int NI = 2 * JT;
      auto issue_item = [&](int n, int slot) {           // one lane: 2 boxes = 2 channel halves of one query row
        mbar_arrive_expect_tx(zbar + 8 * slot, 2048);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          tma_load_3d(zbase + slot * 2048 + q * 1024, &args.tm_z, 32 * q, (n >> 1) * V2_TK,
                      (int)rowb + i0 + r0 + (n & 1), zbar + 8 * slot);
      };
      if (lane == 0)
        for (int n = 0; n < V4_SLOTS && n < NI; ++n) issue_item(n, n);
      int bslot = 0, bphase = 0;                         // slot / phase of the next item the bias phase consumes
      int oslot = 0;                                     // slot of the next item the o_pair phase consumes
      auto next_bias_tile = [&](int par) {
        const int s0 = bslot, p0 = bphase;
        bslot = (bslot == V4_SLOTS - 1) ? 0 : bslot + 1; bphase ^= (bslot == 0);
        const int s1 = bslot, p1 = bphase;
        bslot = (bslot == V4_SLOTS - 1) ? 0 : bslot + 1; bphase ^= (bslot == 0);
        mbar_wait_cta(zbar + 8 * s0, p0);
        mbar_wait_cta(zbar + 8 * s1, p1);
        bias_phase(zgen + s0 * 2048, zgen + s1 * 2048, par);
        if (lane == 0) mbar_arrive(bar_bias + 8 * par);
      };
      next_bias_tile(0);
      for (int jt = 0; jt < JT; ++jt) {
        const int par = jt & 1;
        if (jt + 1 < JT) next_bias_tile(par ^ 1);
        mbar_wait_cta_relaxed(bar_p + 8 * par, (jt >> 1) & 1);   // the head warps have published P / alpha of the tile
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          opair_row(zgen + oslot * 2048, r, par);
          __syncwarp();                                  // every lane is done with the slot
          const int n = 2 * jt + r + V4_SLOTS;
          if (lane == 0 && n < NI) issue_item(n, oslot);
          oslot = (oslot == V4_SLOTS - 1) ? 0 : oslot + 1;
        }
      }
    }
    named_sync(5, V3_THREADS);                         // every warp has left the tile loop: the blob region is free
    float* sop = reinterpret_cast<float*>(smem + LT::SM_BLOB);   // [16 i][8 h][68]: padded rows, conflict-free both ways
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float* o = sop + ((r0 + r) * H + 2 * t) * V2_ZP + 16 * mt + g;
        o[0] = acc[r][mt][0]; o[V2_ZP] = acc[r][mt][1];
        o[8] = acc[r][mt][2]; o[V2_ZP + 8] = acc[r][mt][3];
      }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;\n");
  }
  if constexpr (DEC) tc_fence_before();
  __syncthreads();
  if constexpr (DEC) {
    if (warp == 0) {
      tc_fence_after();
      tmem_dealloc_cta(tmem_base, 256);
    }
  }

  // ---- common epilogue (all 16 warps)
  const float* spt = reinterpret_cast<const float*>(smem + LT::SM_Z);
  const float* sop = reinterpret_cast<const float*>(smem + LT::SM_BLOB);
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
  float* swz = reinterpret_cast<float*>(smem + LT::SM_WDZ);      // [16 d][68]
  for (int idx = tid; idx < 16 * CZ; idx += V3_THREADS) swz[(idx >> 6) * V2_ZP + (idx & 63)] = a.w_dz[idx];
  __syncthreads();
  if (tid < 128) {
    // 4 x 4 register tile per thread: pairs pq + 32 e (pair = i * 8 + head) x outputs dq + 4 e'.  Lanes of a warp read
    // 8 consecutive pair rows / 4 consecutive W_dz rows (pitch 68: conflict-free), so the 16 FMAs of a channel cost
    // 8 shared-memory loads per 4 channels instead of the 20 of a 1 x 4 tile - the epilogue was MIO-bound.
    const int dq = tid & 3, pq = tid >> 2;
    float acc[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { acc[e][0] = acc[e][1] = acc[e][2] = acc[e][3] = 0.f; }
#pragma unroll 2
    for (int c = 0; c < CZ; c += 4) {
      float4 x[4], w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[e] = *reinterpret_cast<const float4*>(sop + (pq + 32 * e) * V2_ZP + c);
        w[e] = *reinterpret_cast<const float4*>(swz + (dq + 4 * e) * V2_ZP + c);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          acc[e][d] = fmaf(w[d].x, x[e].x, acc[e][d]);
          acc[e][d] = fmaf(w[d].y, x[e].y, acc[e][d]);
          acc[e][d] = fmaf(w[d].z, x[e].z, acc[e][d]);
          acc[e][d] = fmaf(w[d].w, x[e].w, acc[e][d]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int pairidx = pq + 32 * e, i = pairidx >> 3, hh = pairidx & 7;
      if (i0 + i >= L) continue;
      const float inv = sl[hh * V2_TQ + i];
      float* out = a.feats + (rowb + i0 + i) * NFEAT + 1024 + 384 + hh * 16;
#pragma unroll
      for (int d = 0; d < 4; ++d) out[dq + 4 * d] = acc[e][d] * inv + a.b_dz[dq + 4 * d];
    }
  }
}

int launch_ipa_attention_v3(const IpaArgs& a, void* workspace, size_t workspace_bytes, cudaStream_t st,
                            bool decoupled) {
  if (a.B == 0 || a.L == 0) return PF_OK;
  PF_REQUIRE(workspace && workspace_bytes >= ipa_v2_workspace_bytes(a.B, a.L), PF_ERR_WORKSPACE_TOO_SMALL);
  const int JT = (a.L + V2_TK - 1) / V2_TK, IT = (a.L + V2_TQ - 1) / V2_TQ;
  unsigned char* blobs = static_cast<unsigned char*>(workspace);
  uint4* Qp = reinterpret_cast<uint4*>(blobs + (size_t)a.B * JT * V2_BLOB);
  IpaPack3Args pa{a.proj, a.rot, a.trans, a.head_w, a.mask, blobs, Qp, a.B, a.L, JT, IT};
  profile_begin(2, st);
  {
    const int nblob = a.B * JT;
    ipa_pack4_kernel<<<nblob < num_sms() ? nblob : num_sms(), P4_THREADS, P4_SMEM, st>>>(pa);
    PF_CHECK_LAUNCH();
    ipa_packq_kernel<<<(unsigned)(a.B * IT), P3_THREADS, PQ_SMEM, st>>>(pa);
  }
  profile_end(2, st);
  PF_CHECK_LAUNCH();
  Ipa3Args args;
  PF_TRY(encode_z_map(&args.tm_z, a.z, a.B, a.L));
  args.p = Ipa2Args{a, blobs, Qp, JT, IT};
  profile_begin(0, st);
  if (decoupled) ipa_attention_v3_kernel<true><<<dim3(IT, a.B), V3_THREADS, V3L<true>::SMEM, st>>>(args);
  else ipa_attention_v3_kernel<false><<<dim3(IT, a.B), V3_THREADS, V3L<false>::SMEM, st>>>(args);
  profile_end(0, st);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

void ipa_v2_kernels_init() {
  cudaFuncSetAttribute(ipa_packq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PQ_SMEM);
  cudaFuncSetAttribute(ipa_pack4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P4_SMEM);
  cudaFuncSetAttribute(ipa_attention_v3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3L<false>::SMEM);
  cudaFuncSetAttribute(ipa_attention_v3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, V3L<true>::SMEM);
}

}  // namespace pf
