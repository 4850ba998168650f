// Cluster launch control (clusterlaunchcontrol.try_cancel) semantics on sm_100a, pinned before the GEMM relies on them:
// grid = one CTA pair per work item; a running pair cancels pending pairs and takes over their item. Checks that every
// item is processed exactly once, reports how many pairs actually ran, with and without another kernel holding SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o test_clc test_clc.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_map(const void* p, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t remote, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(remote), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

constexpr int kSlots = 4;

__global__ void __cluster_dims__(2, 1, 1) clc_kernel(int* hits, int* pairs_ran, int spin) {
  __shared__ alignas(16) uint4 resp[kSlots];
  __shared__ uint64_t full[kSlots], empty[kSlots];
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2 * (blockDim.x / 32) - 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  if (rank == 0 && threadIdx.x == 0) atomicAdd(pairs_ran, 1);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int item = blockIdx.x / 2;
  for (int it = 0;; ++it) {
    const int slot = it % kSlots;
    const uint32_t par = (it / kSlots) & 1;
    if (rank == 0 && warp == 0) {  // scheduler: ask for the next item while this one is processed
      if (lane == 0) {
        mbar_wait(&empty[slot], par ^ 1);
        for (uint32_t r = 0; r < 2; ++r) mbar_expect_tx_cluster(cluster_map(&full[slot], r), 16);
        asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                     ::"r"(smem_u32(&resp[slot])), "r"(smem_u32(&full[slot])) : "memory");
      }
      __syncwarp();
    }
    // ---- "process" the item
    if (threadIdx.x == 0 && rank == 0) atomicAdd(&hits[item], 1);
    for (int s = 0; s < spin; ++s) __nanosleep(1000);
    // ---- next item
    mbar_wait(&full[slot], par);
    uint32_t valid, x;
    asm volatile("{\n.reg .pred p1;\n.reg .b128 r;\nld.shared.b128 r, [%2];\n"
                 "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\nselp.u32 %1, 1, 0, p1;\n"
                 "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n}"
                 : "=r"(x), "=r"(valid) : "r"(smem_u32(&resp[slot])) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0 && !(rank == 0 && warp == 0)) mbar_arrive_cluster(cluster_map(&empty[slot], 0));
    if (!valid) break;
    item = (int)x / 2;
  }
  cluster_sync();
}

__global__ void hog(int us) { for (int i = 0; i < us; ++i) __nanosleep(1000); }

int main() {
  const int items = 3000;
  int *hits, *ran;
  cudaMalloc(&hits, items * sizeof(int)); cudaMalloc(&ran, sizeof(int));
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(hits, 0, items * sizeof(int)); cudaMemset(ran, 0, sizeof(int));
    cudaDeviceSynchronize();
    if (mode == 1) hog<<<8, 1024, 0, s2>>>(20000);   // 8 CTAs of another kernel stay resident for ~20 ms
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s1);
    clc_kernel<<<2 * items, 128, 0, s1>>>(hits, ran, 20);
    cudaEventRecord(e1, s1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    int* h = new int[items]; int r = 0;
    cudaMemcpy(h, hits, items * sizeof(int), cudaMemcpyDeviceToHost); cudaMemcpy(&r, ran, sizeof(int), cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < items; ++i) bad += h[i] != 1;
    printf("mode %d (%s): %s, items processed != once: %d of %d, pairs that ran: %d, %.2f ms (ideal %.2f ms on 74 pairs)\n", mode,
           mode ? "8 SMs held by another kernel" : "alone", cudaGetErrorString(e), bad, items, r, ms, items * 0.020 / 74 * 1.0);
    delete[] h;
  }
  return 0;
}
