// Semantics check of tcgen05.mma.cta_group::2 (bf16, M = 256 over a CTA pair, N = 256, K = 64): each CTA holds its own 128
// rows of A and 128 of the 256 rows of B (K-major, 128-byte swizzle); the leader issues; both CTAs read their 128 x 256
// accumulator back. Prints the number of mismatches against D = A B^T computed on the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I egom2p_b200/csrc tools/micro/test_umma2.cu -o tools/micro/test_umma2 -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
using namespace egom2p;
namespace egom2p { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }

__host__ __device__ inline float a_val(int gr, int k) { return (float)(((gr + 3 * k) % 7) - 3); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((2 * n + k) % 5) - 2); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;            // [128][64] bf16
  uint8_t* sB = smem + 16384;    // [128][64] bf16 : this CTA's half of B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i / 64, kk = i % 64;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((((kk >> 3) ^ (r & 7)) << 4)) + (kk & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = __float2bfloat16(a_val(rank * 128 + r, kk));
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(b_val(rank * 128 + r, kk));
  }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tm = slot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t dA = umma_desc_kmajor_sw128(smem_u32(sA)), dB = umma_desc_kmajor_sw128(smem_u32(sB));
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t acc = kk ? 1u : 0u;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tm), "l"(dA + 2 * kk), "l"(dB + 2 * kk), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  // ---- issue-rate measurement: 512 more pair MMAs on the second accumulator (results unused)
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    const uint64_t dA = umma_desc_kmajor_sw128(smem_u32(sA)), dB = umma_desc_kmajor_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int i = 0; i < 512; ++i) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   ::"r"(tm + 256), "l"(dA + 2 * (i & 3)), "l"(dB + 2 * (i & 3)), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    mbar_wait(&bar, 1);
    out[256 * 256] = (float)(clock64() - t0) / 512.f;
  } else {
    mbar_wait(&bar, 1);
  }
  tc_fence_after();
  for (int c = 0; c < 256; c += 32) {
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[((size_t)(rank * 128 + warp * 32 + lane)) * 256 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "n"(512) : "memory");
  }
}

int main() {
  float* d;
  cudaMalloc(&d, 256 * 256 * 4 + 64);
  cudaMemset(d, 0xff, 256 * 256 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  k<<<2, 128, 40 * 1024>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<float> h(256 * 256 + 1);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  printf("cycles per pair MMA (M=256 over 2 CTAs, N=256, K=16): %.1f\n", h[256 * 256]);
  int bad = 0;
  for (int r = 0; r < 256; ++r)
    for (int n = 0; n < 256; ++n) {
      float ref = 0.f;
      for (int kk = 0; kk < 64; ++kk) ref += a_val(r, kk) * b_val(n, kk);
      if (h[r * 256 + n] != ref) {
        if (bad < 8) printf("mismatch D[%d][%d] = %g, expected %g\n", r, n, h[r * 256 + n], ref);
        ++bad;
      }
    }
  printf("mismatches: %d of 65536\n", bad);
  return 0;
}
