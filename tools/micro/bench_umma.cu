// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, M = 128, K = 16) for the operand layouts the attention kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I egom2p_b200/csrc tools/micro/bench_umma.cu -o tools/micro/bench_umma -lcuda
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
using namespace egom2p;
namespace egom2p { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// mode: 0 = SS K x K, 1 = SS K x MN, 2 = SS MN x MN, 3 = TS x MN, 4 = TS x K, 5 = SS K x K alternating between two
// accumulators, 6 = TS x MN alternating, 7 = SS K x K rotating over 4 accumulators
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int group) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, MODE == 2 ? 1 : 0, (MODE == 1 || MODE == 2 || MODE == 3 || MODE == 6) ? 1 : 0);
    const uint64_t aK = umma_desc_kmajor_sw128(smem_u32(smem)), aM = umma_desc_mnmajor_sw128(smem_u32(smem), 16384);
    const uint64_t bK = umma_desc_kmajor_sw128(smem_u32(smem + 32768)), bM = umma_desc_mnmajor_sw128(smem_u32(smem + 32768), 8192);
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int g = 0; g < group; ++g) {
        const int kk = g & 3;
        if (MODE == 0) umma_bf16_ss(tm, aK + 2 * kk, bK + 2 * kk, idesc, 1u);
        if (MODE == 1) umma_bf16_ss(tm, aK + 2 * kk, bM + 128 * kk, idesc, 1u);
        if (MODE == 2) umma_bf16_ss(tm, aM + 128 * kk, bM + 128 * kk, idesc, 1u);
        if (MODE == 3) umma_ts(tm, tm + 256 + 8 * kk, bM + 128 * kk, idesc, 1u);
        if (MODE == 4) umma_ts(tm, tm + 256 + 8 * kk, bK + 2 * kk, idesc, 1u);
        if (MODE == 5) umma_bf16_ss(tm + (g & 1) * 128, aK + 2 * kk, bK + 2 * kk, idesc, 1u);
        if (MODE == 6) umma_ts(tm + (g & 1) * 128, tm + 256 + 8 * kk, bM + 128 * kk, idesc, 1u);
        if (MODE == 7) umma_bf16_ss(tm + (g & 3) * 64, aK + 2 * kk, bK + 2 * kk, idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

// Two issuing warps (warp 0 and warp 1), each with its own accumulator and mbarrier: is the ~60-cycle floor per issuer or per SM?
template <int N>
__global__ void __launch_bounds__(128, 1) k2(long long* out, int iters, int group, int issuers) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if ((threadIdx.x & 31) == 0 && warp < issuers) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t aK = umma_desc_kmajor_sw128(smem_u32(smem + warp * 16384)), bK = umma_desc_kmajor_sw128(smem_u32(smem + 32768 + warp * 32768));
    long long t0 = clock64();
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int g = 0; g < group; ++g) umma_bf16_ss(tm + warp * 128, aK + 2 * (g & 3), bK + 2 * (g & 3), idesc, 1u);
      umma_commit(&bar[warp]);
      mbar_wait(&bar[warp], phase);
      phase ^= 1;
    }
    long long t1 = clock64();
    if (blockIdx.x == 0 && warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}
template <int N>
void run2(long long* d_out) {
  cudaFuncSetAttribute(k2<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int issuers : {1, 2}) {
    k2<N><<<148, 128, 100 * 1024>>>(d_out, 16, 128, issuers);
    cudaDeviceSynchronize();
    k2<N><<<148, 128, 100 * 1024>>>(d_out, 16, 128, issuers);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("issuers=%d N=%3d: %7.1f cycles per MMA of ONE issuer (%s)\n", issuers, N, (double)h / (16 * 128), cudaGetErrorString(e));
  }
}

template <int N, int MODE>
void run(const char* name, long long* d_out) {
  cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int group : {4, 8, 128}) {
    const int iters = 2048 / group;
    k<N, MODE><<<148, 128, 100 * 1024>>>(d_out, iters, group);
    cudaDeviceSynchronize();
    k<N, MODE><<<148, 128, 100 * 1024>>>(d_out, iters, group);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-22s N=%3d group=%3d: %7.1f cycles / MMA, %7.1f cycles / committed group  (%s)\n", name, N, group,
           (double)h / (iters * group), (double)h / iters, cudaGetErrorString(e));
  }
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  run2<64>(d_out);
  run2<128>(d_out);
  run<64, 5>("SS KxK 2 accumulators", d_out);
  run<128, 5>("SS KxK 2 accumulators", d_out);
  run<64, 6>("TS xMN 2 accumulators", d_out);
  run<64, 7>("SS KxK 4 accumulators", d_out);
  run<64, 0>("SS  K-major x K-major", d_out);
  run<128, 0>("SS  K-major x K-major", d_out);
  run<256, 0>("SS  K-major x K-major", d_out);
  run<64, 1>("SS  K-major x MN-major", d_out);
  run<64, 2>("SS  MN-major x MN-major", d_out);
  run<64, 3>("TS  tmem x MN-major", d_out);
  run<64, 4>("TS  tmem x K-major", d_out);
  run<128, 4>("TS  tmem x K-major", d_out);
  return 0;
}
