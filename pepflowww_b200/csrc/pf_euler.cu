// Manifold maps and the per-iteration update of FlowModel.sample (K10):
//   SO(3) log / exp / geodesic   data/so3_utils.py:88-164,167-282,486-520
//   torus geodesic                models_con/torus.py:5-26
//   denoiser post-processing      models_con/flow_model.py:291-303
//   Euler update                  models_con/flow_model.py:316-333
// One thread per residue; the state is a few dozen floats, so these are coalesced streaming kernels.
#include "pf_common.cuh"

namespace pf {

constexpr float PI_F = 3.14159274101257324f;

// rotvec = Log(R), three branches blended with masks exactly like rotmat_to_rotvec (so3_utils.py:167-254).
__device__ void so3_log_dev(const float* R, float* w) {
  const float vx = R[7] - R[5], vy = R[2] - R[6], vz = R[3] - R[1];
  const float sin_t = sqrtf(vx * vx + vy * vy + vz * vz) / 2.0f;
  const float cos_t = (R[0] + R[4] + R[8] - 1.0f) / 2.0f;
  const float th = atan2f(sin_t, cos_t);
  const float m0 = (fabsf(th) <= 1e-8f) ? 1.f : 0.f;                        // isclose(th, 0)
  const float mpi = (fabsf(th - PI_F) <= (1e-2f + 1e-5f * PI_F)) ? 1.f : 0.f;  // isclose(th, pi, atol=1e-2)
  const float me = (1.f - m0) * (1.f - mpi);
  const float num = m0 / 2.0f + th * me;
  const float den = (1.0f - th * th / 6.0f) * m0 + 2.0f * sin_t * me + mpi;
  const float pref = num / den;
  w[0] = vx * pref; w[1] = vy * pref; w[2] = vz * pref;
  if (mpi != 0.f) {
    // omega omega^T = (I + R) / 2 with the diagonal clamped at 0; signs from the row of largest norm
    float M[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = (((e == 0 || e == 4 || e == 8) ? 1.0f : 0.0f) + R[e]) / 2.0f;
    M[0] = fmaxf(M[0], 0.f); M[4] = fmaxf(M[4], 0.f); M[8] = fmaxf(M[8], 0.f);
    const float n0 = sqrtf(M[0] * M[0] + M[1] * M[1] + M[2] * M[2]);
    const float n1 = sqrtf(M[3] * M[3] + M[4] * M[4] + M[5] * M[5]);
    const float n2 = sqrtf(M[6] * M[6] + M[7] * M[7] + M[8] * M[8]);
    int r = 0;
    float best = n0;
    if (n1 > best) { best = n1; r = 1; }
    if (n2 > best) { best = n2; r = 2; }
    const float d[3] = {sqrtf(M[0]), sqrtf(M[4]), sqrtf(M[8])};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float sel = (r == 0) ? M[k] : ((r == 1) ? M[3 + k] : M[6 + k]);
      const float sg = (sel > 0.f) ? 1.f : ((sel < 0.f) ? -1.f : 0.f);
      w[k] += d[k] * th * sg;
    }
  }
}

// R = Exp(w): Rodrigues with the Taylor fallback below 1e-7 (so3_utils.py:88-164).
__device__ void so3_exp_dev(const float* w, float* R) {
  const float th = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const float th2 = th * th;
  float s, c;
  if (fabsf(th) < 1e-7f) {
    s = 1.0f - th2 / 6.0f;
    c = 0.5f - th2 / 24.0f;
  } else {
    s = sinf(th) / th;
    c = (1.0f - cosf(th)) / th2;
  }
  const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float kk = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
      R[i * 3 + j] = ((i == j) ? 1.0f : 0.0f) + s * K[i * 3 + j] + c * kk;
    }
}

// out = base * Exp(t * Log(base^T mat))   (so3_utils.py:486-520)
// __noinline__: every caller (so3_geodesic_kernel, euler_step_kernel, sampler_update_kernel) runs the SAME instruction
// sequence, so the one-call iteration and the three-call iteration agree bit for bit (no per-call-site FMA contraction)
__device__ __noinline__ void so3_geodesic_dev(float t, const float* mat, const float* base, float* out) {
  float rel[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) rel[i * 3 + j] = base[i] * mat[j] + base[3 + i] * mat[3 + j] + base[6 + i] * mat[6 + j];
  float w[3];
  so3_log_dev(rel, w);
  w[0] *= t; w[1] *= t; w[2] *= t;
  float E[9];
  so3_exp_dev(w, E);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      out[i * 3 + j] = base[i * 3] * E[j] + base[i * 3 + 1] * E[3 + j] + base[i * 3 + 2] * E[6 + j];
}

__device__ __noinline__ float tor_geodesic_dev(float t, float a1, float a0) {
  const float d = a1 - a0;
  return mod_2pi(a0 + t * atan2f(sinf(d), cosf(d)));
}

// Philox4x32-10 (Salmon et al. 2011): counter-based, one call gives four 32-bit words.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float uniform_for(const float* uniforms, uint64_t seed, uint64_t counter, int n) {
  if (uniforms) return uniforms[n];
  uint32_t o[4];
  philox4x32_10((uint32_t)n, (uint32_t)counter, (uint32_t)(counter >> 32), 0x5eedu, (uint32_t)seed,
                (uint32_t)(seed >> 32), o);
  return (float)(o[0] >> 8) * (1.0f / 16777216.0f);
}

// categorical draw from softmax(x[0..19]) + 1e-8 by inverse CDF (stands in for multinomial,
// pepflow/modules/common/layers.py:17-22); same left-to-right fp32 prefix sums as the oracle.
__device__ __noinline__ int categorical20(const float* x, float u) {
  float mx = x[0];
#pragma unroll
  for (int k = 1; k < 20; ++k) mx = fmaxf(mx, x[k]);
  float p[20], s = 0.f;
#pragma unroll
  for (int k = 0; k < 20; ++k) { p[k] = expf(x[k] - mx); s += p[k]; }
  float total = 0.f;
#pragma unroll
  for (int k = 0; k < 20; ++k) { p[k] = p[k] / s + 1e-8f; total += p[k]; }
  const float thr = u * total;
  float run = 0.f;
  int idx = 0;
#pragma unroll
  for (int k = 0; k < 19; ++k) { run += p[k]; idx += (run <= thr) ? 1 : 0; }
  return idx;
}

__global__ void so3_log_kernel(const float* __restrict__ rot, float* __restrict__ w, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float R[9], o[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) R[e] = rot[(size_t)i * 9 + e];
  so3_log_dev(R, o);
  w[(size_t)i * 3] = o[0]; w[(size_t)i * 3 + 1] = o[1]; w[(size_t)i * 3 + 2] = o[2];
}
__global__ void so3_exp_kernel(const float* __restrict__ w, float* __restrict__ rot, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v[3] = {w[(size_t)i * 3], w[(size_t)i * 3 + 1], w[(size_t)i * 3 + 2]};
  float R[9];
  so3_exp_dev(v, R);
#pragma unroll
  for (int e = 0; e < 9; ++e) rot[(size_t)i * 9 + e] = R[e];
}
__global__ void so3_geodesic_kernel(const float* __restrict__ t, const float* __restrict__ mat,
                                    const float* __restrict__ base, float* __restrict__ out, int n, int group) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float A[9], Bm[9], O[9];
#pragma unroll
  for (int e = 0; e < 9; ++e) { A[e] = mat[(size_t)i * 9 + e]; Bm[e] = base[(size_t)i * 9 + e]; }
  so3_geodesic_dev(t[i / group], A, Bm, O);
#pragma unroll
  for (int e = 0; e < 9; ++e) out[(size_t)i * 9 + e] = O[e];
}
__global__ void tor_geodesic_kernel(const float* __restrict__ t, const float* __restrict__ a1,
                                    const float* __restrict__ a0, float* __restrict__ out, int n, int group, int d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d) return;
  out[i] = tor_geodesic_dev(t[(i / d) / group], a1[i], a0[i]);
}

struct PostArgs {
  const float* pred_rot; const float* pred_trans; const float* pred_ang; const float* logits;
  const float* rot1; const float* trans1; const float* ang1; const int64_t* seq1; const uint8_t* gen;
  const float* tmask; const float* uniforms; uint64_t seed, counter;
  float* c_rot; float* c_trans; float* c_ang; int64_t* c_seq; float* c_simplex;
  int n; float k;
};
__global__ void denoise_post_kernel(PostArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const bool g = a.gen[i] != 0;
#pragma unroll
  for (int e = 0; e < 9; ++e) a.c_rot[(size_t)i * 9 + e] = g ? a.pred_rot[(size_t)i * 9 + e] : a.rot1[(size_t)i * 9 + e];
#pragma unroll
  for (int e = 0; e < 3; ++e) a.c_trans[(size_t)i * 3 + e] = g ? a.pred_trans[(size_t)i * 3 + e] : a.trans1[(size_t)i * 3 + e];
  int64_t s = a.seq1[i];
  if (g) {
    float lg[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) lg[k] = a.logits[(size_t)i * 20 + k];
    s = categorical20(lg, uniform_for(a.uniforms, a.seed, a.counter, i));
  }
  a.c_seq[i] = s;
  const int si = (s >= 0 && s < 22) ? (int)s : 21;
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const float v = g ? a.pred_ang[(size_t)i * 5 + e] : a.ang1[(size_t)i * 5 + e];
    a.c_ang[(size_t)i * 5 + e] = (a.tmask[si * 5 + e] != 0.f) ? v : 0.f;
  }
#pragma unroll
  for (int k = 0; k < 20; ++k) a.c_simplex[(size_t)i * 20 + k] = (s == k) ? a.k : -a.k;  // one_hot*2k - k
}

struct EulerArgs {
  const float* rot_t; const float* trans_t; const float* ang_t; const float* simplex_t;
  const float* c_rot; const float* c_trans; const float* c_ang; const int64_t* c_seq;
  const float* trans0; const float* simplex0; const float* rot1; const float* trans1; const float* ang1;
  const int64_t* seq1; const uint8_t* gen; const float* tmask; const float* uniforms; uint64_t seed, counter;
  float d_t;
  float* rot_o; float* trans_o; float* ang_o; int64_t* seq_o; float* simplex_o;
  int n; float k;
};
__global__ void euler_step_kernel(EulerArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  const bool g = a.gen[i] != 0;
  const float dt = a.d_t;
  // translations (flow_model.py:318-320): x_t + (x^ - x_0) d_t   -- x_0 is the initial noise
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const size_t o = (size_t)i * 3 + e;
    const float v = __fadd_rn(a.trans_t[o], __fmul_rn(a.c_trans[o] - a.trans0[o], dt));   // two roundings, like torch
    a.trans_o[o] = g ? v : a.trans1[o];
  }
  // rotations (:322-323): geodesic with the fixed 10 d_t schedule
  {
    float Rt[9], Rh[9], O[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) { Rt[e] = a.rot_t[(size_t)i * 9 + e]; Rh[e] = a.c_rot[(size_t)i * 9 + e]; }
    if (g) so3_geodesic_dev(dt * 10.0f, Rh, Rt, O);
#pragma unroll
    for (int e = 0; e < 9; ++e) a.rot_o[(size_t)i * 9 + e] = g ? O[e] : a.rot1[(size_t)i * 9 + e];
  }
  // simplex + residue types (:328-330); no where() on the simplex
  float sx[20];
  const int64_t sh = a.c_seq[i];
#pragma unroll
  for (int k = 0; k < 20; ++k) {
    const size_t o = (size_t)i * 20 + k;
    const float target = (sh == k) ? a.k : -a.k;
    sx[k] = __fadd_rn(a.simplex_t[o], __fmul_rn(target - a.simplex0[o], dt));
    a.simplex_o[o] = sx[k];
  }
  int64_t s2 = a.seq1[i];
  if (g) s2 = categorical20(sx, uniform_for(a.uniforms, a.seed, a.counter, i));
  a.seq_o[i] = s2;
  const int si = (s2 >= 0 && s2 < 22) ? (int)s2 : 21;
  // torsions (:325-326, :332-333)
#pragma unroll
  for (int e = 0; e < 5; ++e) {
    const size_t o = (size_t)i * 5 + e;
    const float v = g ? tor_geodesic_dev(dt, a.c_ang[o], a.ang_t[o]) : a.ang1[o];
    a.ang_o[o] = (a.tmask[si * 5 + e] != 0.f) ? v : 0.f;
  }
}


// ---- one whole loop iteration with device-resident bookkeeping (pf_sampler_step) -----------------------------------
// sampler_begin_kernel: n = step[0] is the iteration this call executes: t_cur[b] = ts[n], step[1] = n, step[0] = n + 1.
__global__ void sampler_begin_kernel(int* __restrict__ step, const float* __restrict__ ts, float* __restrict__ t_cur,
                                     int B, int num_steps) {
  int n = step[0];
  n = n < 0 ? 0 : (n > num_steps - 1 ? num_steps - 1 : n);
  const float t = ts[n];
  for (int b = threadIdx.x; b < B; b += blockDim.x) t_cur[b] = t;
  __syncthreads();
  if (threadIdx.x == 0) { step[1] = n; step[0] = n + 1; }
}

constexpr int SU_T = 128;            // residues per CTA
constexpr int SU_PITCH = 21;         // staging pitch of the 20-wide rows (odd: conflict-free row-per-thread access)

struct SamplerArgs {
  const float* pred_rot; const float* pred_trans; const float* pred_ang; const float* logits;
  const float* rot1; const float* trans1; const float* ang1; const int64_t* seq1; const uint8_t* gen;
  const float* tmask; const float* trans0; const float* simplex0;
  float* rot_t; float* trans_t; float* ang_t; int64_t* seq_t; float* simplex_t;
  float* traj_rot; float* traj_trans; float* traj_ang; int64_t* traj_seq; float* traj_simplex;
  const float* ts; const float* uniforms; const int* step;
  uint64_t seed;
  int n, num_steps;                  // n = B * L residues
  int sample_bb, sample_ang, sample_seq;
  float k;
};

// Rows of W floats cross HBM as contiguous, fully coalesced runs of the CTA's 128-residue tile (staged in shared
// memory); every thread then owns one residue in registers.
template <int W, int P>
__device__ __forceinline__ void su_load(float* stage, const float* __restrict__ g, int r0, int nr, float (&v)[W]) {
  __syncthreads();
  const float* src = g + (size_t)r0 * W;
  for (int e = threadIdx.x; e < nr * W; e += SU_T) stage[(e / W) * P + (e % W)] = src[e];
  __syncthreads();
  if ((int)threadIdx.x < nr) {
#pragma unroll
    for (int e = 0; e < W; ++e) v[e] = stage[threadIdx.x * P + e];
  }
}
template <int W, int P>
__device__ __forceinline__ void su_store(float* stage, float* __restrict__ g, int r0, int nr, const float (&v)[W]) {
  __syncthreads();
  if ((int)threadIdx.x < nr) {
#pragma unroll
    for (int e = 0; e < W; ++e) stage[threadIdx.x * P + e] = v[e];
  }
  __syncthreads();
  float* dst = g + (size_t)r0 * W;
  for (int e = threadIdx.x; e < nr * W; e += SU_T) dst[e] = stage[(e / W) * P + (e % W)];
}

// Post-processing of the denoiser output into trajectory slot n (flow_model.py:291-314) and the Euler update of the
// state (:316-343), one residue per thread.
__global__ void __launch_bounds__(SU_T) sampler_update_kernel(SamplerArgs a) {
  __shared__ float stage[SU_T * SU_PITCH];
  __shared__ float s_tmask[22 * 5];
  const int r0 = blockIdx.x * SU_T;
  const int nr = min(SU_T, a.n - r0);
  const int i = r0 + threadIdx.x;
  const bool on = (int)threadIdx.x < nr;
  if (threadIdx.x < 110) s_tmask[threadIdx.x] = a.tmask[threadIdx.x];
  const int n = a.step[1];
  const size_t slot = (size_t)n * a.n;            // residue offset of trajectory slot n
  const bool g = on && a.gen[i] != 0;
  const int64_t s1 = on ? a.seq1[i] : 0;

  float c_rot[9], c_trans[3], c_ang[5], rot1[9], trans1[3], ang1[5];
  su_load<9, 9>(stage, a.rot1, r0, nr, rot1);
  su_load<3, 3>(stage, a.trans1, r0, nr, trans1);
  su_load<5, 5>(stage, a.ang1, r0, nr, ang1);
  su_load<9, 9>(stage, a.pred_rot, r0, nr, c_rot);
  su_load<3, 3>(stage, a.pred_trans, r0, nr, c_trans);
  su_load<5, 5>(stage, a.pred_ang, r0, nr, c_ang);
  int64_t c_seq = s1;
  {
    float lg[20];
    su_load<20, SU_PITCH>(stage, a.logits, r0, nr, lg);
    if (g) {
      const float* u = a.uniforms ? a.uniforms + (size_t)(2 * n) * a.n : nullptr;
      c_seq = categorical20(lg, uniform_for(u, a.seed, (uint64_t)(2 * n), i));
    }
  }
  if (!g || !a.sample_bb) {
#pragma unroll
    for (int e = 0; e < 9; ++e) c_rot[e] = rot1[e];
#pragma unroll
    for (int e = 0; e < 3; ++e) c_trans[e] = trans1[e];
  }
  {
    const int si = (c_seq >= 0 && c_seq < 22) ? (int)c_seq : 21;
#pragma unroll
    for (int e = 0; e < 5; ++e) c_ang[e] = (s_tmask[si * 5 + e] != 0.f) ? (g ? c_ang[e] : ang1[e]) : 0.f;
  }
  if (!a.sample_ang) {
#pragma unroll
    for (int e = 0; e < 5; ++e) c_ang[e] = ang1[e];
  }
  if (!a.sample_seq) c_seq = s1;
  su_store<9, 9>(stage, a.traj_rot + slot * 9, r0, nr, c_rot);
  su_store<3, 3>(stage, a.traj_trans + slot * 3, r0, nr, c_trans);
  su_store<5, 5>(stage, a.traj_ang + slot * 5, r0, nr, c_ang);
  if (on) a.traj_seq[slot + i] = c_seq;
  {
    float cs[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) cs[k] = (c_seq == k) ? a.k : -a.k;      // one_hot * 2k - k (flow_model.py:108-109)
    su_store<20, SU_PITCH>(stage, a.traj_simplex + slot * 20, r0, nr, cs);
  }
  if (n >= a.num_steps - 1) return;               // the last iteration only records its prediction (:346-372)

  const float dt = a.ts[n + 1] - a.ts[n];
  // translations (:318-320): x_t + (x^ - x_0) d_t   -- x_0 is the initial noise
  {
    float xt[3], x0[3];
    su_load<3, 3>(stage, a.trans_t, r0, nr, xt);
    su_load<3, 3>(stage, a.trans0, r0, nr, x0);
#pragma unroll
    for (int e = 0; e < 3; ++e) xt[e] = (g && a.sample_bb) ? __fadd_rn(xt[e], __fmul_rn(c_trans[e] - x0[e], dt)) : trans1[e];
    su_store<3, 3>(stage, a.trans_t, r0, nr, xt);
  }
  // rotations (:322-323): geodesic with the fixed 10 d_t schedule
  {
    float Rt[9], O[9];
    su_load<9, 9>(stage, a.rot_t, r0, nr, Rt);
    const bool live = g && a.sample_bb;
    if (live) so3_geodesic_dev(dt * 10.0f, c_rot, Rt, O);
#pragma unroll
    for (int e = 0; e < 9; ++e) O[e] = live ? O[e] : rot1[e];
    su_store<9, 9>(stage, a.rot_t, r0, nr, O);
  }
  // simplex + residue types (:328-330); no where() on the simplex
  int64_t s2 = s1;
  {
    float sx[20], s0[20];
    su_load<20, SU_PITCH>(stage, a.simplex_t, r0, nr, sx);
    su_load<20, SU_PITCH>(stage, a.simplex0, r0, nr, s0);
#pragma unroll
    for (int k = 0; k < 20; ++k) sx[k] = __fadd_rn(sx[k], __fmul_rn(((c_seq == k) ? a.k : -a.k) - s0[k], dt));
    su_store<20, SU_PITCH>(stage, a.simplex_t, r0, nr, sx);
    if (g) {
      const float* u = a.uniforms ? a.uniforms + (size_t)(2 * n + 1) * a.n : nullptr;
      s2 = categorical20(sx, uniform_for(u, a.seed, (uint64_t)(2 * n + 1), i));
    }
  }
  // torsions (:325-326, :332-333)
  {
    float at[5];
    su_load<5, 5>(stage, a.ang_t, r0, nr, at);
    const int si = (s2 >= 0 && s2 < 22) ? (int)s2 : 21;
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const float v = g ? tor_geodesic_dev(dt, c_ang[e], at[e]) : ang1[e];
      at[e] = a.sample_ang ? ((s_tmask[si * 5 + e] != 0.f) ? v : 0.f) : ang1[e];
    }
    su_store<5, 5>(stage, a.ang_t, r0, nr, at);
  }
  if (on) a.seq_t[i] = a.sample_seq ? s2 : s1;
}

// FlowModel.zero_center_part (flow_model.py:95-106): one CTA per complex
__global__ void __launch_bounds__(256) zero_center_kernel(float* __restrict__ pos, const uint8_t* __restrict__ gen,
                                                           const float* __restrict__ rmask, float* __restrict__ center,
                                                           int L) {
  __shared__ float red[4][8];
  __shared__ float c[3];
  const int b = blockIdx.x, tid = threadIdx.x;
  float* p = pos + (size_t)b * L * 3;
  float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
  for (int l = tid; l < L; l += 256) {
    const float m = gen[(size_t)b * L + l] ? 1.f : 0.f;
    sx += p[l * 3] * m; sy += p[l * 3 + 1] * m; sz += p[l * 3 + 2] * m; cnt += m;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); cnt = warp_sum(cnt);
  if ((tid & 31) == 0) { red[0][tid >> 5] = sx; red[1][tid >> 5] = sy; red[2][tid >> 5] = sz; red[3][tid >> 5] = cnt; }
  __syncthreads();
  if (tid < 3) {
    float s = 0.f, n = 0.f;
    for (int w = 0; w < 8; ++w) { s += red[tid][w]; n += red[3][w]; }
    c[tid] = s / (n + 1e-8f);
    if (center) center[b * 3 + tid] = c[tid];
  }
  __syncthreads();
  for (int e = tid; e < L * 3; e += 256) p[e] = (p[e] - c[e % 3]) * rmask[(size_t)b * L + e / 3];
}

}  // namespace pf

extern "C" {

int pf_so3_log(const float* rot, float* rotvec, int n, void* stream) {
  PF_REQUIRE(rot && rotvec, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::so3_log_kernel<<<(n + 127) / 128, 128, 0, pf::as_stream(stream)>>>(rot, rotvec, n);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
int pf_so3_exp(const float* rotvec, float* rot, int n, void* stream) {
  PF_REQUIRE(rot && rotvec, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::so3_exp_kernel<<<(n + 127) / 128, 128, 0, pf::as_stream(stream)>>>(rotvec, rot, n);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
int pf_so3_geodesic(const float* t, const float* mat, const float* base, float* out, int n, int group, void* stream) {
  PF_REQUIRE(t && mat && base && out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0 && group > 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::so3_geodesic_kernel<<<(n + 127) / 128, 128, 0, pf::as_stream(stream)>>>(t, mat, base, out, n, group);
  PF_CHECK_LAUNCH();
  return PF_OK;
}
int pf_tor_geodesic(const float* t, const float* ang1, const float* ang0, float* out, int n, int group, int d,
                    void* stream) {
  PF_REQUIRE(t && ang1 && ang0 && out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0 && group > 0 && d > 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::tor_geodesic_kernel<<<(n * d + 255) / 256, 256, 0, pf::as_stream(stream)>>>(t, ang1, ang0, out, n, group, d);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int pf_denoise_post(const float* pred_rot, const float* pred_trans, const float* pred_ang, const float* logits,
                    const float* rot1, const float* trans1, const float* ang1, const int64_t* seq1,
                    const uint8_t* gen_mask, const float* torsions_mask, const float* uniforms, uint64_t seed,
                    uint64_t counter, float* clean_rot, float* clean_trans, float* clean_ang, int64_t* clean_seq,
                    float* clean_simplex, int n, float simplex_k, void* stream) {
  PF_REQUIRE(pred_rot && pred_trans && pred_ang && logits && rot1 && trans1 && ang1 && seq1 && gen_mask &&
                 torsions_mask && clean_rot && clean_trans && clean_ang && clean_seq && clean_simplex,
             PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::PostArgs a{pred_rot, pred_trans, pred_ang, logits, rot1, trans1, ang1, seq1, gen_mask, torsions_mask, uniforms,
                 seed, counter, clean_rot, clean_trans, clean_ang, clean_seq, clean_simplex, n, simplex_k};
  pf::denoise_post_kernel<<<(n + 127) / 128, 128, 0, pf::as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int pf_euler_step(const float* rot_t, const float* trans_t, const float* ang_t, const float* simplex_t,
                  const float* clean_rot, const float* clean_trans, const float* clean_ang, const int64_t* clean_seq,
                  const float* trans0, const float* simplex0, const float* rot1, const float* trans1,
                  const float* ang1, const int64_t* seq1, const uint8_t* gen_mask, const float* torsions_mask,
                  const float* uniforms, uint64_t seed, uint64_t counter, float d_t, float* rot_out, float* trans_out,
                  float* ang_out, int64_t* seq_out, float* simplex_out, int n, float simplex_k, void* stream) {
  PF_REQUIRE(rot_t && trans_t && ang_t && simplex_t && clean_rot && clean_trans && clean_ang && clean_seq && trans0 &&
                 simplex0 && rot1 && trans1 && ang1 && seq1 && gen_mask && torsions_mask && rot_out && trans_out &&
                 ang_out && seq_out && simplex_out, PF_ERR_NULL_POINTER);
  PF_REQUIRE(n >= 0, PF_ERR_BAD_SHAPE);
  if (n == 0) return PF_OK;
  pf::EulerArgs a{rot_t, trans_t, ang_t, simplex_t, clean_rot, clean_trans, clean_ang, clean_seq, trans0, simplex0,
                  rot1, trans1, ang1, seq1, gen_mask, torsions_mask, uniforms, seed, counter, d_t,
                  rot_out, trans_out, ang_out, seq_out, simplex_out, n, simplex_k};
  pf::euler_step_kernel<<<(n + 127) / 128, 128, 0, pf::as_stream(stream)>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int pf_sampler_step(const pf_sampler* s, void* stream) {
  PF_REQUIRE(s && s->weights && s->node_embed && s->edge_embed && s->res_mask && s->workspace && s->rot1 && s->trans1 &&
                 s->ang1 && s->seq1 && s->gen_mask && s->torsions_mask && s->trans0 && s->simplex0 && s->rot_t &&
                 s->trans_t && s->ang_t && s->seq_t && s->simplex_t && s->pred_rot && s->pred_trans && s->pred_ang &&
                 s->logits && s->traj_rot && s->traj_trans && s->traj_ang && s->traj_seq && s->traj_simplex && s->ts &&
                 s->step && s->t_cur, PF_ERR_NULL_POINTER);
  PF_REQUIRE(s->B >= 0 && s->L >= 0 && s->num_steps >= 1, PF_ERR_BAD_SHAPE);
  if (s->B == 0 || s->L == 0) return PF_OK;
  cudaStream_t st = pf::as_stream(stream);
  pf::sampler_begin_kernel<<<1, 128, 0, st>>>(s->step, s->ts, s->t_cur, s->B, s->num_steps);
  PF_CHECK_LAUNCH();
  PF_TRY(pf_ga_encoder_forward(s->weights, s->t_cur, s->rot_t, s->trans_t, s->ang_t, s->seq_t, s->node_embed,
                               s->edge_embed, s->res_mask, s->pred_rot, s->pred_trans, s->pred_ang, s->logits, nullptr,
                               s->workspace, (size_t)s->workspace_bytes, s->B, s->L, stream));
  const int n = s->B * s->L;
  pf::SamplerArgs a{s->pred_rot, s->pred_trans, s->pred_ang, s->logits, s->rot1, s->trans1, s->ang1, s->seq1,
                    s->gen_mask, s->torsions_mask, s->trans0, s->simplex0, s->rot_t, s->trans_t, s->ang_t, s->seq_t,
                    s->simplex_t, s->traj_rot, s->traj_trans, s->traj_ang, s->traj_seq, s->traj_simplex, s->ts,
                    s->uniforms, s->step, s->seed, n, s->num_steps, s->sample_bb, s->sample_ang, s->sample_seq,
                    s->simplex_k};
  pf::sampler_update_kernel<<<(n + pf::SU_T - 1) / pf::SU_T, pf::SU_T, 0, st>>>(a);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

int pf_zero_center(float* pos, const uint8_t* gen_mask, const float* res_mask, float* center_out, int B, int L,
                   void* stream) {
  PF_REQUIRE(pos && gen_mask && res_mask, PF_ERR_NULL_POINTER);
  PF_REQUIRE(B >= 0 && L >= 0, PF_ERR_BAD_SHAPE);
  if (B == 0 || L == 0) return PF_OK;
  pf::zero_center_kernel<<<B, 256, 0, pf::as_stream(stream)>>>(pos, gen_mask, res_mask, center_out, L);
  PF_CHECK_LAUNCH();
  return PF_OK;
}

}  // extern "C"
