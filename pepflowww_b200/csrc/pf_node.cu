// Per-residue (O(L)) kernels of the denoiser: input feature mix, residual+LayerNorm, sequence
// transformer attention core, rigid-frame update, IPA point projection into the global frame.
#include "pf_common.cuh"
#include "pf_split.cuh"

namespace pf {

// ---------------------------------------------------------------- K1: feature mix input
// x[row, 0:629] = node_embed | seq_emb[seq] | [sin,cos](t*2056*w_i) | per angle [a, sin(a f), cos(a f)]
// (models_con/ga.py:94, models_con/utils.py:60-72, pepflow/modules/common/layers.py:104-113)
__global__ void mix_features_kernel(const float* __restrict__ node, const float* __restrict__ emb,
                                    const int64_t* __restrict__ seqs, const float* __restrict__ t,
                                    const float* __restrict__ tfreq, const float* __restrict__ angles,
                                    const float* __restrict__ afreq, float* __restrict__ x, int B, int L, int ldx) {
  // ldx = NMIX, or NMIX rounded up to a multiple of 128 with zero fill (the K loop of the fused layer chain)
  const size_t total = (size_t)B * L * ldx;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldx);
    const size_t row = idx / ldx;
    float v;
    if (col >= NMIX) {
      v = 0.f;
    } else if (col < 128) {
      v = node[row * 128 + col];
    } else if (col < 256) {
      v = emb[seqs[row] * 128 + (col - 128)];
    } else if (col < 384) {
      const int i = col - 256;
      const float tau = t[row / L] * 2056.0f;
      const float arg = tau * tfreq[i & 63];
      v = (i < 64) ? sinf(arg) : cosf(arg);
    } else {
      const int a = col - 384;
      const int ai = a / 49, r = a % 49;
      const float ang = angles[row * 5 + ai];
      if (r == 0) v = ang;
      else if (r <= 24) v = sinf(ang * afreq[r - 1]);
      else v = cosf(ang * afreq[r - 25]);
    }
    x[idx] = v;
  }
}

// ---------------------------------------------------------------- residual + LayerNorm (+ row mask)
// one warp per row, N = 32 * VPL
template <int VPL>
__global__ void add_layernorm_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     const float* __restrict__ rowmask, float* __restrict__ y, int M) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  constexpr int N = 32 * VPL;
  float v[VPL];
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < VPL; ++e) {
    const int c = lane + 32 * e;
    v[e] = a[(size_t)warp * N + c] + (b ? b[(size_t)warp * N + c] : 0.f);
    s += v[e];
  }
  const float mu = warp_sum(s) / N;
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < VPL; ++e) {
    const float d = v[e] - mu;
    q += d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / N + 1e-5f);
  const float rm = rowmask ? rowmask[warp] : 1.f;
#pragma unroll
  for (int e = 0; e < VPL; ++e) {
    const int c = lane + 32 * e;
    y[(size_t)warp * N + c] = ((v[e] - mu) * rstd * gamma[c] + beta[c]) * rm;
  }
}

// ---------------------------------------------------------------- K5: transformer attention core
// CTA = (head, complex); K/V of this head staged in smem; warp handles query rows i = warp, warp+8, ...
// qkv row layout (torch in_proj): q(128) | k(128) | v(128), head h at columns h*32 .. h*32+31.
//
// Flash-style on the tensor cores (mma.sync m16n8k16, 3xFP16 split precision, fp32 accumulate): CTA = one
// (complex, head); K and V^T of the head are staged ONCE in shared memory as fp16 hi / lo halves in
// bank-conflict-free B-fragment order; the 9 warps walk the 16-row query tiles of the complex (stride 9) and stream
// the keys in chunks of 32 with an online softmax, re-using the score fragments as the A operand of P V.
// (One CTA per 128 query rows re-staged K / V three times at L = 271: 61 us per call.)
constexpr int TFH = 4;
constexpr int SEQ_WARPS = 9;     // 288 threads, 2 CTAs per SM
constexpr int SEQ_KW = 20;       // words (half2) per key row of K: 16 + 4 pad  -> conflict-free fragment loads

__host__ __device__ inline int seq_lp(int L) { return (L + 31) & ~31; }
__host__ __device__ inline int seq_vw(int L) { return seq_lp(L) / 2 + 12; }   // words per d row of V^T

__global__ void __launch_bounds__(SEQ_WARPS * 32, 2) seq_attention_kernel(const float* __restrict__ qkv,
                                                            const float* __restrict__ mask,
                                                            float* __restrict__ ctx, int L) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  const int Lp = seq_lp(L), VW = seq_vw(L);
  uint32_t* Kh = smem_u;                   // [Lp][SEQ_KW]   half2 (c, c+1)
  uint32_t* Kl = Kh + (size_t)Lp * SEQ_KW;
  uint32_t* Vh = Kl + (size_t)Lp * SEQ_KW; // [32 d][VW]     half2 (key, key+1)
  uint32_t* Vl = Vh + (size_t)32 * VW;
  float* Ms = reinterpret_cast<float*>(Vl + (size_t)32 * VW);   // [Lp] additive key mask: 0 or -inf
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (size_t)b * L * 384;

  // ---- stage K (row-major pairs along c) and V^T (pairs along the key index)
  constexpr int NT = SEQ_WARPS * 32;
  for (int idx = tid; idx < Lp * 16; idx += NT) {
    const int j = idx >> 4, w = idx & 15;
    float2 k2 = make_float2(0.f, 0.f);
    if (j < L) k2 = *reinterpret_cast<const float2*>(base + (size_t)j * 384 + 128 + h * 32 + 2 * w);
    uint32_t hi, lo;
    split_pair(k2.x, k2.y, hi, lo);
    Kh[j * SEQ_KW + w] = hi;
    Kl[j * SEQ_KW + w] = lo;
  }
  for (int idx = tid; idx < (Lp / 2) * 32; idx += NT) {
    const int d = idx & 31, jp = idx >> 5, j = 2 * jp;
    const float v0 = (j < L) ? base[(size_t)j * 384 + 256 + h * 32 + d] : 0.f;
    const float v1 = (j + 1 < L) ? base[(size_t)(j + 1) * 384 + 256 + h * 32 + d] : 0.f;
    uint32_t hi, lo;
    split_pair(v0, v1, hi, lo);
    Vh[d * VW + jp] = hi;
    Vl[d * VW + jp] = lo;
  }
  for (int j = tid; j < Lp; j += NT) Ms[j] = (j < L && mask[(size_t)b * L + j] != 0.f) ? 0.f : -INFINITY;

  __syncthreads();
  for (int i0 = warp * 16; i0 < L; i0 += SEQ_WARPS * 16) {
  // ---- Q fragments of this warp's 16 rows (1/sqrt(32) folded in)
  const int i_lo = i0 + g, i_hi = i0 + g + 8;
  uint32_t qh[2][4], ql[2][4];
  {
    const float scale = 0.17677669529663687f;  // 1/sqrt(32)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + 2 * t;
      float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      if (i_lo < L) {
        a0 = *reinterpret_cast<const float2*>(base + (size_t)i_lo * 384 + h * 32 + c);
        a2 = *reinterpret_cast<const float2*>(base + (size_t)i_lo * 384 + h * 32 + c + 8);
      }
      if (i_hi < L) {
        a1 = *reinterpret_cast<const float2*>(base + (size_t)i_hi * 384 + h * 32 + c);
        a3 = *reinterpret_cast<const float2*>(base + (size_t)i_hi * 384 + h * 32 + c + 8);
      }
      split_pair(a0.x * scale, a0.y * scale, qh[ks][0], ql[ks][0]);
      split_pair(a1.x * scale, a1.y * scale, qh[ks][1], ql[ks][1]);
      split_pair(a2.x * scale, a2.y * scale, qh[ks][2], ql[ks][2]);
      split_pair(a3.x * scale, a3.y * scale, qh[ks][3], ql[ks][3]);
    }
  }

  float O[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n) { O[n][0] = O[n][1] = O[n][2] = O[n][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

  for (int j0 = 0; j0 < Lp; j0 += 32) {
    // S = Q K^T for 32 keys (4 n-tiles of 8)
    float S[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
      const int key = j0 + nt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t bh0 = Kh[key * SEQ_KW + ks * 8 + t], bh1 = Kh[key * SEQ_KW + ks * 8 + t + 4];
        const uint32_t bl0 = Kl[key * SEQ_KW + ks * 8 + t], bl1 = Kl[key * SEQ_KW + ks * 8 + t + 4];
        mma16816(S[nt], ql[ks], bh0, bh1);
        mma16816(S[nt], qh[ks], bl0, bl1);
        mma16816(S[nt], qh[ks], bh0, bh1);
      }
    }
    // key-padding mask, online softmax (rows g and g+8; a row lives in the 4 lanes of a quad)
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float2 mk = *reinterpret_cast<const float2*>(Ms + j0 + nt * 8 + 2 * t);
      S[nt][0] += mk.x; S[nt][1] += mk.y; S[nt][2] += mk.x; S[nt][3] += mk.y;
      mx_lo = fmaxf(mx_lo, fmaxf(S[nt][0], S[nt][1]));
      mx_hi = fmaxf(mx_hi, fmaxf(S[nt][2], S[nt][3]));
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);
    // rows whose keys so far are all masked keep m = -inf: use 0 as the reference point (every p becomes 0)
    const float rf_lo = (mn_lo == -INFINITY) ? 0.f : mn_lo, rf_hi = (mn_hi == -INFINITY) ? 0.f : mn_hi;
    const float al_lo = __expf(m_lo - rf_lo), al_hi = __expf(m_hi - rf_hi);   // ex2.approx: ~3e-6 relative over a row
    m_lo = mn_lo; m_hi = mn_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      S[nt][0] = __expf(S[nt][0] - rf_lo); S[nt][1] = __expf(S[nt][1] - rf_lo);
      S[nt][2] = __expf(S[nt][2] - rf_hi); S[nt][3] = __expf(S[nt][3] - rf_hi);
      ps_lo += S[nt][0] + S[nt][1];
      ps_hi += S[nt][2] + S[nt][3];
    }
    l_lo = l_lo * al_lo + ps_lo;
    l_hi = l_hi * al_hi + ps_hi;
#pragma unroll
    for (int n = 0; n < 4; ++n) { O[n][0] *= al_lo; O[n][1] *= al_lo; O[n][2] *= al_hi; O[n][3] *= al_hi; }
    // O += P V: the score fragments of n-tiles (2k, 2k+1) are the A fragment of key step k
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ph[4], pl[4];
      split_pair(S[2 * ks][0], S[2 * ks][1], ph[0], pl[0]);
      split_pair(S[2 * ks][2], S[2 * ks][3], ph[1], pl[1]);
      split_pair(S[2 * ks + 1][0], S[2 * ks + 1][1], ph[2], pl[2]);
      split_pair(S[2 * ks + 1][2], S[2 * ks + 1][3], ph[3], pl[3]);
      const int w0 = (j0 >> 1) + ks * 8 + t;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int d = n * 8 + g;
        const uint32_t bh0 = Vh[d * VW + w0], bh1 = Vh[d * VW + w0 + 4];
        const uint32_t bl0 = Vl[d * VW + w0], bl1 = Vl[d * VW + w0 + 4];
        mma16816(O[n], pl, bh0, bh1);
        mma16816(O[n], ph, bl0, bl1);
        mma16816(O[n], ph, bh0, bh1);
      }
    }
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float il_lo = l_lo > 0.f ? 1.0f / l_lo : 0.f, il_hi = l_hi > 0.f ? 1.0f / l_hi : 0.f;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int d = n * 8 + 2 * t;
    if (i_lo < L)
      *reinterpret_cast<float2*>(ctx + ((size_t)b * L + i_lo) * 128 + h * 32 + d) = make_float2(O[n][0] * il_lo, O[n][1] * il_lo);
    if (i_hi < L)
      *reinterpret_cast<float2*>(ctx + ((size_t)b * L + i_hi) * 128 + h * 32 + d) = make_float2(O[n][2] * il_hi, O[n][3] * il_hi);
  }
  }   // row tiles
}

// ---------------------------------------------------------------- K7: rigid update
// rot -> quat at block 0: Shepperd's closed form as the start vector, refined by three power
// iterations on K(R)/3 + I/3 whose dominant eigenvector is the reference's eigh answer
// (openfold/utils/rigid_utils.py:208-227); the sign is free (R(q) and the update are sign-invariant).
__device__ void rot_to_quat_dev(const float* R, float* q) {
  const float xx = R[0], xy = R[1], xz = R[2], yx = R[3], yy = R[4], yz = R[5], zx = R[6], zy = R[7], zz = R[8];
  float k[4][4] = {{xx + yy + zz, zy - yz, xz - zx, yx - xy},
                   {zy - yz, xx - yy - zz, xy + yx, xz + zx},
                   {xz - zx, xy + yx, yy - xx - zz, yz + zy},
                   {yx - xy, xz + zx, yz + zy, zz - xx - yy}};
  // Start from the column of M = K/3 + I/3 with the largest diagonal (for an exact rotation that
  // column is already (4/3) q_best q, and max_i q_i^2 >= 1/4 keeps it away from zero).
  int best = 0;
  float bd = k[0][0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (k[i][i] > bd) { bd = k[i][i]; best = i; }
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = k[i][best] / 3.0f + (i == best ? 1.0f / 3.0f : 0.0f);
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    float n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
    const float inv = rsqrtf(n2);
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = v[i] * inv;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] = (k[i][0] * u[0] + k[i][1] * u[1] + k[i][2] * u[2] + k[i][3] * u[3]) / 3.0f + u[i] / 3.0f;
  }
  const float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = v[i] / n;
}

// q' = normalize(q + m * q (x) (0,v));  t' = t + m * R_old u   (rigid_utils.py:1039-1063,587-616,266-275)
__global__ void rigid_update_kernel(const float* __restrict__ quat_in, const float* __restrict__ rot_in,
                                    const float* __restrict__ trans_in, const float* __restrict__ upd,
                                    const float* __restrict__ mask, float* __restrict__ quat_out,
                                    float* __restrict__ rot_out, float* __restrict__ trans_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float q[4], R[9];
  if (quat_in) {
#pragma unroll
    for (int e = 0; e < 4; ++e) q[e] = quat_in[(size_t)i * 4 + e];
    quat_to_rot_dev(q[0], q[1], q[2], q[3], R);
  } else {
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = rot_in[(size_t)i * 9 + e];
    rot_to_quat_dev(R, q);
  }
  const float m = mask[i];
  const float* u = upd + (size_t)i * 6;
  const float x = u[0], y = u[1], z = u[2];
  const float a = q[0], b = q[1], c = q[2], d = q[3];
  float nq[4] = {a + m * (-b * x - c * y - d * z), b + m * (a * x + c * z - d * y),
                 c + m * (a * y - b * z + d * x), d + m * (a * z + b * y - c * x)};
  const float nn = sqrtf(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2] + nq[3] * nq[3]);
#pragma unroll
  for (int e = 0; e < 4; ++e) nq[e] /= nn;
  const float tx = u[3], ty = u[4], tz = u[5];
  trans_out[(size_t)i * 3 + 0] = trans_in[(size_t)i * 3 + 0] + m * (R[0] * tx + R[1] * ty + R[2] * tz);
  trans_out[(size_t)i * 3 + 1] = trans_in[(size_t)i * 3 + 1] + m * (R[3] * tx + R[4] * ty + R[5] * tz);
  trans_out[(size_t)i * 3 + 2] = trans_in[(size_t)i * 3 + 2] + m * (R[6] * tx + R[7] * ty + R[8] * tz);
  float Rn[9];
  quat_to_rot_dev(nq[0], nq[1], nq[2], nq[3], Rn);
#pragma unroll
  for (int e = 0; e < 4; ++e) quat_out[(size_t)i * 4 + e] = nq[e];
#pragma unroll
  for (int e = 0; e < 9; ++e) rot_out[(size_t)i * 9 + e] = Rn[e];
}

__global__ void quat_to_rot_kernel(const float* __restrict__ quat, float* __restrict__ rot, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9];
  quat_to_rot_dev(quat[(size_t)i * 4], quat[(size_t)i * 4 + 1], quat[(size_t)i * 4 + 2], quat[(size_t)i * 4 + 3], R);
#pragma unroll
  for (int e = 0; e < 9; ++e) rot[(size_t)i * 9 + e] = R[e];
}

__global__ void mod_2pi_kernel(const float* __restrict__ x, float* __restrict__ y, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = mod_2pi(x[i]);
}

// ---------------------------------------------------------------- K2 epilogue: points -> global frame
// proj columns: q_pts at OFF_QP as [3][H][8], kv_pts at OFF_KVP as [3][H][20] (x|y|z planes, head-major,
// models_con/ipa_pytorch.py:360-387).  pts[row][h][28][3] = R * local + t  (rigid_utils.py:1124-1136).
__global__ void ipa_points_kernel(const float* __restrict__ proj, const float* __restrict__ rot,
                                  const float* __restrict__ trans, float* __restrict__ pts, int rows) {
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)rows * H * NPT;
  if (idx >= total) return;
  const int p = (int)(idx % NPT);
  const int h = (int)((idx / NPT) % H);
  const size_t row = idx / (NPT * H);
  const float* pr = proj + row * NPROJ;
  float lx, ly, lz;
  if (p < PQ) {
    const int o = OFF_QP + h * PQ + p;
    lx = pr[o]; ly = pr[o + H * PQ]; lz = pr[o + 2 * H * PQ];
  } else {
    const int n = PQ + PV;
    const int o = OFF_KVP + h * n + (p - PQ);
    lx = pr[o]; ly = pr[o + H * n]; lz = pr[o + 2 * H * n];
  }
  const float* R = rot + row * 9;
  const float* t = trans + row * 3;
  float* out = pts + idx * 3;
  out[0] = R[0] * lx + R[1] * ly + R[2] * lz + t[0];
  out[1] = R[3] * lx + R[4] * ly + R[5] * lz + t[1];
  out[2] = R[6] * lx + R[7] * ly + R[8] * lz + t[2];
}

// ---------------------------------------------------------------- launchers (internal)
int launch_mix_features(const float* node, const float* emb, const int64_t* seqs, const float* t, const float* tfreq,
                        const float* angles, const float* afreq, float* x, int B, int L, cudaStream_t st, int ldx) {
  const size_t total = (size_t)B * L * ldx;
  if (total == 0) return PF_OK;
  const int blocks = (int)min((size_t)(num_sms() * 8), (total + 255) / 256);
  mix_features_kernel<<<blocks, 256, 0, st>>>(node, emb, seqs, t, tfreq, angles, afreq, x, B, L, ldx);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, const float* rowmask,
                         float* y, int M, int N, cudaStream_t st) {
  if (M == 0) return PF_OK;
  const int blocks = (M + 7) / 8;
  if (N == 128) add_layernorm_kernel<4><<<blocks, 256, 0, st>>>(a, b, gamma, beta, rowmask, y, M);
  else if (N == 64) add_layernorm_kernel<2><<<blocks, 256, 0, st>>>(a, b, gamma, beta, rowmask, y, M);
  else return PF_ERR_BAD_SHAPE;
  PF_CHECK_LAUNCH();
  return PF_OK;
}

size_t seq_attention_smem(int L) {
  return ((size_t)seq_lp(L) * SEQ_KW * 2 + (size_t)32 * seq_vw(L) * 2 + seq_lp(L)) * sizeof(uint32_t);
}

int launch_seq_attention(const float* qkv, const float* mask, float* ctx, int B, int L, cudaStream_t st) {
  if (B == 0 || L == 0) return PF_OK;
  const size_t smem = seq_attention_smem(L);
  if (smem > 227 * 1024) return PF_ERR_BAD_SHAPE;  // L <= ~980
  seq_attention_kernel<<<dim3(TFH, B), SEQ_WARPS * 32, smem, st>>>(qkv, mask, ctx, L);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_rigid_update(const float* quat_in, const float* rot_in, const float* trans_in, const float* upd,
                        const float* mask, float* quat_out, float* rot_out, float* trans_out, int n, cudaStream_t st) {
  if (n == 0) return PF_OK;
  rigid_update_kernel<<<(n + 127) / 128, 128, 0, st>>>(quat_in, rot_in, trans_in, upd, mask, quat_out, rot_out,
                                                       trans_out, n);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_ipa_points(const float* proj, const float* rot, const float* trans, float* pts, int rows, cudaStream_t st) {
  const size_t total = (size_t)rows * H * NPT;
  if (total == 0) return PF_OK;
  ipa_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(proj, rot, trans, pts, rows);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_mod_2pi(const float* x, float* y, int n, cudaStream_t st) {
  if (n == 0) return PF_OK;
  mod_2pi_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, y, n);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int launch_quat_to_rot(const float* quat, float* rot, int n, cudaStream_t st) {
  if (n == 0) return PF_OK;
  quat_to_rot_kernel<<<(n + 255) / 256, 256, 0, st>>>(quat, rot, n);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

void node_kernels_init() {
  cudaFuncSetAttribute(seq_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

}  // namespace pf

// ---------------------------------------------------------------- C ABI
extern "C" {

int pf_add_layernorm(const float* a, const float* b, const float* gamma, const float* beta, const float* rowmask,
                     float* y, int M, int N, void* stream) {
  PF_REQUIRE(a && gamma && beta && y, PF_ERR_NULL_POINTER);
  PF_REQUIRE(M >= 0, PF_ERR_BAD_SHAPE);
  return pf::launch_add_layernorm(a, b, gamma, beta, rowmask, y, M, N, pf::as_stream(stream));
}

int pf_mix_features(const float* node_embed, const float* seq_emb_table, const int64_t* seqs, const float* t,
                    const float* time_freqs, const float* angles, const float* ang_freqs, float* x, int B, int L,
                    void* stream) {
  PF_REQUIRE(node_embed && seq_emb_table && seqs && t && time_freqs && angles && ang_freqs && x, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  return pf::launch_mix_features(node_embed, seq_emb_table, seqs, t, time_freqs, angles, ang_freqs, x, B, L,
                                 pf::as_stream(stream));
}

int pf_ipa_points(const float* proj, const float* rot, const float* trans, float* pts, int B, int L, void* stream) {
  PF_REQUIRE(proj && rot && trans && pts, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  return pf::launch_ipa_points(proj, rot, trans, pts, B * L, pf::as_stream(stream));
}

int pf_seq_attention(const float* qkv, const float* mask, float* ctx, int B, int L, void* stream) {
  PF_REQUIRE(qkv && mask && ctx, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  return pf::launch_seq_attention(qkv, mask, ctx, B, L, pf::as_stream(stream));
}

int pf_rigid_update(const float* quat_in, const float* rot_in, const float* trans_in, const float* upd,
                    const float* mask, float* quat_out, float* rot_out, float* trans_out, int n, void* stream) {
  PF_REQUIRE((quat_in || rot_in) && trans_in && upd && mask && quat_out && rot_out && trans_out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0, PF_ERR_BAD_SHAPE);
  return pf::launch_rigid_update(quat_in, rot_in, trans_in, upd, mask, quat_out, rot_out, trans_out, n,
                                 pf::as_stream(stream));
}

int pf_mod_2pi(const float* x, float* y, int n, void* stream) {
  PF_REQUIRE(x && y, PF_ERR_NULL_POINTER);
  return pf::launch_mod_2pi(x, y, n, pf::as_stream(stream));
}

int pf_quat_to_rot(const float* quat, float* rot, int n, void* stream) {
  PF_REQUIRE(quat && rot, PF_ERR_NULL_POINTER);
  return pf::launch_quat_to_rot(quat, rot, n, pf::as_stream(stream));
}

}  // extern "C"
