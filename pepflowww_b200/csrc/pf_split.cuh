// Split-precision helpers for the tensor-core kernels ("3xFP16"): an fp32 value x is carried as
//   hi = fp16(x),  lo = fp16(x - hi)         (x - hi is exact in fp32)
// and a product is accumulated in fp32 as  hi*hi' + lo*hi' + hi*lo'  (the lo*lo' term, ~2^-24, is dropped).
// fp16 keeps 11 significand bits per half, so the pair carries ~22 bits: GEMM error ~1e-7 relative, i.e.
// fp32-grade - measured on B200: a 3xBF16 split (8+8 bits) left 1e-5 per GEMM and 7.7e-5 on the rotations
// after six blocks, too close to the 1e-4 parity bar; single-pass TF32/BF16 fail it outright
// (SURVEY.md finding 5).  Conversions saturate at +-65504 instead of producing inf.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace pf {

// packed {lower 16 bits = fp16(x0), upper 16 bits = fp16(x1)}, round-to-nearest, saturating
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float x0, float x1) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(x1), "f"(x0));
  return d;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
// (x0, x1) -> packed hi pair and packed lo pair
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_f16x2_sat(x0, x1);
  const float2 h = unpack_f16x2(hi);
  lo = cvt_f16x2_sat(x0 - h.x, x1 - h.y);
}
__device__ __forceinline__ void split_one(float x, __half& hi, __half& lo) {
  const uint32_t h = cvt_f16x2_sat(x, 0.f);
  hi = __ushort_as_half(static_cast<unsigned short>(h & 0xffffu));
  const uint32_t l = cvt_f16x2_sat(x - __half2float(hi), 0.f);
  lo = __ushort_as_half(static_cast<unsigned short>(l & 0xffffu));
}

// D(16x8, fp32) += A(16x16, fp16, row) * B(16x8, fp16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace pf
