// Device-side input / target masking of the token modalities.
//
// Reference: UnifiedMasking.image_mask (egom2p/data/masking.py:236-266), run per sample and modality in the data-loader
// workers: noise = torch.rand(L); ids_shuffle = argsort(noise); the input_budget positions that come first in that random
// order are encoder inputs (input_mask False), the next target_budget ones are decoder targets (target_mask False), and
// decoder_attention_mask carries the target count at the first target position. Here one CTA does that for one sample of
// one modality: 32-bit Philox draws (or caller-supplied noise, for parity tests) become 64-bit keys (value, position) --
// unique, so the order is a permutation and equal to a stable argsort -- which are sorted with a bitonic network in
// shared memory; ids_shuffle[r] = key[r].position then decides the two masks of position r exactly as the reference's
// gather does. The (B, L) masks never exist on the host.
#include <curand_kernel.h>

#include "common.cuh"

namespace egom2p {

constexpr int kMaskThreads = 1024;
constexpr int kMaskMaxLen = 8192;

struct MaskParams {
  int B, L, n2;
  const float* noise;           // (B, L) or NULL
  unsigned long long seed, offset;
  int stream_id;                // Philox subsequence base (modality index), so that modalities draw independent streams
  const int32_t* input_budget;  // (B)
  const int32_t* target_budget; // (B) or NULL (targets = everything that is not an input)
  uint8_t* input_mask;          // (B, L) 1 = masked
  uint8_t* target_mask;         // (B, L)
  int32_t* attn_cnt;            // (B, L)
};

__global__ void __launch_bounds__(kMaskThreads) image_mask_kernel(MaskParams p) {
  extern __shared__ unsigned long long keys[];
  __shared__ int s_first;
  const int b = blockIdx.x, t = threadIdx.x;
  curandStatePhilox4_32_10_t st;
  if (!p.noise) curand_init(p.seed, (unsigned long long)p.stream_id * p.B * kMaskThreads + (unsigned long long)b * kMaskThreads + t, p.offset, &st);
  for (int i = t; i < p.n2; i += kMaskThreads) {
    unsigned long long k = ~0ull;
    if (i < p.L) {
      const uint32_t v = p.noise ? __float_as_uint(p.noise[(int64_t)b * p.L + i]) : curand(&st);
      k = ((unsigned long long)v << 32) | (uint32_t)i;
    }
    keys[i] = k;
  }
  if (t == 0) s_first = 0x7fffffff;
  __syncthreads();
  // bitonic sort, ascending
  for (int size = 2; size <= p.n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = t; i < (p.n2 >> 1); i += kMaskThreads) {
        const int lo = 2 * i - (i & (stride - 1));   // index with the `stride` bit clear
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  const int ib = min(max(p.input_budget[b], 0), p.L);
  const int tb = p.target_budget ? min(max(p.target_budget[b], 0), p.L - ib) : p.L - ib;
  // the reference gathers [0] * budget + [1] * rest THROUGH ids_shuffle (masking.py:252-260): position r is an input iff
  // the element of rank r has an index below the budget
  for (int r = t; r < p.L; r += kMaskThreads) {
    const int idx = (int)(keys[r] & 0xffffffffu);   // ids_shuffle[r]
    const bool is_in = idx < ib, is_tg = idx >= ib && idx < ib + tb;
    p.input_mask[(int64_t)b * p.L + r] = is_in ? 0 : 1;
    p.target_mask[(int64_t)b * p.L + r] = is_tg ? 0 : 1;
    p.attn_cnt[(int64_t)b * p.L + r] = 0;
    if (is_tg) atomicMin(&s_first, r);
  }
  __syncthreads();
  if (t == 0 && tb > 0) p.attn_cnt[(int64_t)b * p.L + s_first] = tb;
}

}  // namespace egom2p

extern "C" int egom2p_image_masks(const float* noise, uint64_t seed, uint64_t offset, int32_t stream_id, int32_t B, int32_t L,
                                  const int32_t* input_budget, const int32_t* target_budget, uint8_t* input_mask,
                                  uint8_t* target_mask, int32_t* attn_cnt, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(input_budget && input_mask && target_mask && attn_cnt && B > 0 && L > 0, "image_masks: bad argument");
  EGO_REQUIRE(L <= kMaskMaxLen, "image_masks: at most %d tokens per modality (got %d)", kMaskMaxLen, L);
  int n2 = 32;
  while (n2 < L) n2 <<= 1;
  MaskParams p{B, L, n2, noise, seed, offset, stream_id, input_budget, target_budget, input_mask, target_mask, attn_cnt};
  static std::atomic<uint64_t> attr_done{0};
  int rc = ensure_dyn_smem(image_mask_kernel, kMaskMaxLen * 8, attr_done, "image_masks");
  if (rc) return rc;
  image_mask_kernel<<<B, kMaskThreads, (size_t)n2 * 8, (cudaStream_t)stream>>>(p);
  return check_launch("image_masks");
}
