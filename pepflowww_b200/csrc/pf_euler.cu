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
__device__ void so3_geodesic_dev(float t, const float* mat, const float* base, float* out) {
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

__device__ __forceinline__ float tor_geodesic_dev(float t, float a1, float a0) {
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
__device__ int categorical20(const float* x, float u) {
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
    const float v = a.trans_t[o] + (a.c_trans[o] - a.trans0[o]) * dt;
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
    sx[k] = a.simplex_t[o] + (target - a.simplex0[o]) * dt;
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

}  // extern "C"
