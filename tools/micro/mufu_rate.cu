// Micro-benchmark: MUFU.EX2 throughput per SM vs warps per SM sub-partition, alone and mixed with the softmax's other
// instructions (FFMA2 + F2FP + HADD2).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 1e-3f + i * 0.01f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ex2(v[i]);       // 32 independent MUFU per iteration
    } else {
      // softmax-like: per pair: 1 FFMA2-equivalent (2 FFMA), 2 MUFU, 1 F2FP, ~0.75 HADD2
      unsigned w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = fmaf(v[2 * i], 0.125f, -1.0f), b = fmaf(v[2 * i + 1], 0.125f, -1.0f);
        const float e0 = ex2(a), e1 = ex2(b);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(e1), "f"(e0));
      }
      __half2 acc = *reinterpret_cast<__half2*>(&w[0]);
#pragma unroll
      for (int i = 1; i < 16; ++i) acc = __hadd2(acc, *reinterpret_cast<__half2*>(&w[i]));
      const float2 f = __half22float2(acc);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] * 0.5f + (i & 1 ? f.x : f.y) * 1e-6f;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {4, 8, 12, 16, 32}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
      }
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double mufu = double(iters) * 32 * warps * 32;   // per SM
      printf("mode %d warps/SM %2d (%d per SMSP): %lld clk, %.2f MUFU/clk/SM\n", mode, warps, warps / 4, c, mufu / c);
    }
  return 0;
}
