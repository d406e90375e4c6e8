// Micro-benchmark: throughput of 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
template <int NACC>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x * 1e-7f, y = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
  u64 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo = threadIdx.x * 1e-3f + i, hi = lo + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(lo), "f"(hi));
  }
  u64 x, y;
  float xa = a + threadIdx.x * 1e-7f;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(xa), "f"(xa));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b), "f"(b));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma2(acc[i], x, y);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int ctas = p.multiProcessorCount * 8, thr = 256, iters = 4096;
  float* out;
  cudaMalloc(&out, sizeof(float) * ctas * thr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](auto kern, const char* name, double fma_per_inst, int nacc) {
    kern<<<ctas, thr>>>(out, 0.999f, 1e-3f, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<ctas, thr>>>(out, 0.999f, 1e-3f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)ctas * thr * iters * nacc * fma_per_inst;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz)\n", name, ms, 2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  };
  run(k_ffma<16>, "FFMA  3-reg, 16 acc", 1, 16);
  run(k_ffma<32>, "FFMA  3-reg, 32 acc", 1, 32);
  run(k_ffma2<16>, "FFMA2 packed, 16 acc pairs", 2, 16);
  run(k_ffma2<32>, "FFMA2 packed, 32 acc pairs", 2, 32);
  return 0;
}
