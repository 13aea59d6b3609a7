// Small geometry device functions shared by the once-per-sample kernels (pf_embed.cu, pf_recon.cu).
#pragma once
#include <cuda_runtime.h>

namespace pf {

// Signed dihedral of four points, NaN for degenerate geometry (models_con/torsion.py:13-29, _get_torsion).
__device__ __forceinline__ float dihedral4_raw(const float* p0, const float* p1, const float* p2, const float* p3) {
  const float v0x = p2[0] - p1[0], v0y = p2[1] - p1[1], v0z = p2[2] - p1[2];
  const float v1x = p0[0] - p1[0], v1y = p0[1] - p1[1], v1z = p0[2] - p1[2];
  const float v2x = p3[0] - p2[0], v2y = p3[1] - p2[1], v2z = p3[2] - p2[2];
  const float u1x = v0y * v1z - v0z * v1y, u1y = v0z * v1x - v0x * v1z, u1z = v0x * v1y - v0y * v1x;
  const float u2x = v0y * v2z - v0z * v2y, u2y = v0z * v2x - v0x * v2z, u2z = v0x * v2y - v0y * v2x;
  const float l1 = sqrtf(u1x * u1x + u1y * u1y + u1z * u1z), l2 = sqrtf(u2x * u2x + u2y * u2y + u2z * u2z);
  const float d = (u1x / l1) * (u2x / l2) + (u1y / l1) * (u2y / l2) + (u1z / l1) * (u2z / l2);
  if (!(d == d)) return d;                         // fminf / fmaxf would swallow the NaN
  const float cx = v1y * v2z - v1z * v2y, cy = v1z * v2x - v1x * v2z, cz = v1x * v2y - v1y * v2x;
  const float s = cx * v0x + cy * v0y + cz * v0z;
  const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
  return sg * acosf(fminf(fmaxf(d, -0.999999f), 0.999999f));
}

// dihedral_from_four_points (pepflow/modules/common/geometry.py:296-313) with nan_to_num, ROUNDING FOR ROUNDING like the
// reference's PyTorch CPU kernels.  acos(clamp(cos, +-0.999999)) amplifies an input rounding difference up to 700 x and
// the sign comes from a triple product that is ~0 for planar atoms, so a merely "fp32-accurate" cosine leaves 1e-3
// differences in the pair embedding of ~1000 pairs per complex, which the denoiser amplifies past the 1e-4 bar.  The
// sequence below reproduces torch.cross / linalg.norm / sum bit for bit (checked against torch 2.11 CPU, AVX2 / AVX512
// builds, on every tensor layout the reference uses; tests/test_host_logic.py::test_dihedral_rounding_model):
//   cross:  c = fma(a1, b2, -(a2 * b1))           norm: sqrt(fma(z, z, fma(y, y, x * x)))
//   u / norm: IEEE division                        dot / triple product: ((m0 + m1) + m2) of rounded products
__device__ __forceinline__ float cross_c(float a1, float b2, float a2, float b1) {
  return __fmaf_rn(a1, b2, -__fmul_rn(a2, b1));
}
__device__ __forceinline__ float dot3_lr(float ax, float ay, float az, float bx, float by, float bz) {
  return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}
__device__ __forceinline__ float dihedral4(const float* p0, const float* p1, const float* p2, const float* p3) {
  const float v0x = __fsub_rn(p2[0], p1[0]), v0y = __fsub_rn(p2[1], p1[1]), v0z = __fsub_rn(p2[2], p1[2]);
  const float v1x = __fsub_rn(p0[0], p1[0]), v1y = __fsub_rn(p0[1], p1[1]), v1z = __fsub_rn(p0[2], p1[2]);
  const float v2x = __fsub_rn(p3[0], p2[0]), v2y = __fsub_rn(p3[1], p2[1]), v2z = __fsub_rn(p3[2], p2[2]);
  const float u1x = cross_c(v0y, v1z, v0z, v1y), u1y = cross_c(v0z, v1x, v0x, v1z), u1z = cross_c(v0x, v1y, v0y, v1x);
  const float u2x = cross_c(v0y, v2z, v0z, v2y), u2y = cross_c(v0z, v2x, v0x, v2z), u2z = cross_c(v0x, v2y, v0y, v2x);
  const float l1 = __fsqrt_rn(__fmaf_rn(u1z, u1z, __fmaf_rn(u1y, u1y, __fmul_rn(u1x, u1x))));
  const float l2 = __fsqrt_rn(__fmaf_rn(u2z, u2z, __fmaf_rn(u2y, u2y, __fmul_rn(u2x, u2x))));
  const float d = dot3_lr(__fdiv_rn(u1x, l1), __fdiv_rn(u1y, l1), __fdiv_rn(u1z, l1),
                          __fdiv_rn(u2x, l2), __fdiv_rn(u2y, l2), __fdiv_rn(u2z, l2));
  if (!(d == d)) return 0.f;                       // degenerate geometry (padded residues): nan_to_num
  const float cx = cross_c(v1y, v2z, v1z, v2y), cy = cross_c(v1z, v2x, v1x, v2z), cz = cross_c(v1x, v2y, v1y, v2x);
  const float s = dot3_lr(cx, cy, cz, v0x, v0y, v0z);
  const float sg = (s > 0.f) ? 1.f : ((s < 0.f) ? -1.f : 0.f);
  const float a = sg * acosf(fminf(fmaxf(d, -0.999999f), 0.999999f));
  return a == a ? a : 0.f;
}

// C = A B for row-major 3x3 matrices
__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = fmaf(A[i * 3 + 2], B[6 + j], fmaf(A[i * 3 + 1], B[3 + j], A[i * 3] * B[j]));
}
// y = R x + t
__device__ __forceinline__ void rigid_apply(const float* R, const float* t, const float* x, float* y) {
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = fmaf(R[i * 3 + 2], x[2], fmaf(R[i * 3 + 1], x[1], R[i * 3] * x[0])) + t[i];
}

}  // namespace pf
