// Microbenchmark: MUFU throughput of ex2.approx.ftz.f32 against the packed ex2.approx.ftz.bf16x2 / f16x2 forms on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/bench_ex2.cu -o tools/micro/bench_ex2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xbc00bc00u + threadIdx.x + i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  if (s == 12345.678f) out[0] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
  float* out;
  cudaMalloc(&out, 4);
  const int iters = 4096;
  k<MODE><<<148, 1024>>>(out, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ops = 148.0 * 1024 * 8.0 * iters;          // MUFU lane-instructions
  printf("%-26s %8.3f ms  %6.1f G lane-ops/s  %6.1f G exponentials/s  (%.1f lane-ops / clk / SM at the max clock %d MHz)  err=%d\n",
         name, ms, ops / ms / 1e6, ops * per_op / ms / 1e6, ops / (ms * 1e-3) / 148.0 / (khz * 1e3), khz / 1000, (int)cudaGetLastError());
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  return 0;
}
