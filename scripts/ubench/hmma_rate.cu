// Micro-benchmark (diagnostic): issue rate of legacy mma.sync shapes on sm_100a, per SM sub-core.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/hmma_rate scripts/ubench/hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int SHAPE, int CHAINS>
__global__ void k(float* out, long long* cyc, int iters) {
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  uint32_t a0 = threadIdx.x * 0x3c003c00u, a1 = 0x3c003c00u, a2 = 0x38003800u, a3 = 0x3c003800u, b0 = 0x3c003c00u, b1 = 0x34003400u;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (SHAPE == 16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int SHAPE, int CHAINS>
void run(int warps, float* out, long long* cyc) {
  const int iters = 2000;
  k<SHAPE, CHAINS><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_subcore = (double)iters * CHAINS * warps / 4.0;   // MMAs issued on one sub-core
  printf("m16n8k%-2d chains %d warps/SM %2d: %.2f cycles per MMA per sub-core (%.1f cycles per MMA per warp)\n", SHAPE, CHAINS,
         warps, (double)c / per_subcore, (double)c / (iters * CHAINS));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16}) {
    run<16, 1>(warps, out, cyc); run<16, 4>(warps, out, cyc); run<16, 8>(warps, out, cyc);
    run<8, 1>(warps, out, cyc); run<8, 4>(warps, out, cyc); run<8, 8>(warps, out, cyc);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
