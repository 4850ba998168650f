// Cluster launch control without clusters (cluster size 1, 3-D grid): every (x, y, z) processed exactly once.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o test_clc1 test_clc1.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(128) k(int* hits, int* ran, int gx, int gy) {
  extern __shared__ uint8_t dyn[];   // large dynamic smem: one CTA per SM, like the attention backward
  __shared__ alignas(16) uint4 resp;
  __shared__ uint64_t full;
  if (threadIdx.x == 0) { mbar_init(&full, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); atomicAdd(ran, 1); }
  __syncthreads();
  int x = blockIdx.x, y = blockIdx.y, z = blockIdx.z;
  for (int it = 0;; ++it) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(smem_u32(&full)) : "memory");
      asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                   ::"r"(smem_u32(&resp)), "r"(smem_u32(&full)) : "memory");
      atomicAdd(&hits[(z * gy + y) * gx + x], 1);
    }
    for (int s = 0; s < 20; ++s) __nanosleep(500);
    mbar_wait(&full, it & 1);
    uint32_t valid, nx, ny, nz;
    asm volatile("{\n.reg .pred p1;\n.reg .b128 r;\nld.shared.b128 r, [%4];\n"
                 "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\nselp.u32 %3, 1, 0, p1;\n"
                 "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0;\n"
                 "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, %1, %2, _}, r;\n}"
                 : "=r"(nx), "=r"(ny), "=r"(nz), "=r"(valid) : "r"(smem_u32(&resp)) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();   // everybody has read the response before thread 0 asks again
    if (!valid) break;
    x = nx; y = ny; z = nz;
  }
}
int main() {
  const int gx = 16, gy = 12, gz = 16, items = gx * gy * gz;
  int *hits, *ran;
  cudaMalloc(&hits, items * sizeof(int)); cudaMalloc(&ran, sizeof(int));
  cudaMemset(hits, 0, items * sizeof(int)); cudaMemset(ran, 0, sizeof(int));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<dim3(gx, gy, gz), 128, 200 * 1024>>>(hits, ran, gx, gy);
  cudaError_t e = cudaDeviceSynchronize();
  int* h = new int[items]; int r = 0;
  cudaMemcpy(h, hits, items * sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(&r, ran, sizeof(int), cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < items; ++i) bad += h[i] != 1;
  printf("%s: items processed != once: %d of %d, CTAs that ran: %d\n", cudaGetErrorString(e), bad, items, r);
  return bad != 0;
}
