// Shared pieces of the attention kernels: tiling constants, range metadata layout, small device helpers.
#pragma once
#include <climits>

#include "common.cuh"

namespace egom2p {

constexpr int kAttnComputeWarps = 8;                       // two warps per TMEM lane quarter (column halves)
constexpr int kAttnThreads = 32 * (kAttnComputeWarps + 2); // + TMA warp + MMA warp
constexpr int kTmaWarp = kAttnComputeWarps, kMmaWarp = kAttnComputeWarps + 1;
constexpr int kD = 64;     // head dim
constexpr int kT = 128;    // tile rows (queries in fwd / dQ, keys in dKV)
constexpr int kBlk = 64;   // inner block (keys in fwd / dQ, query rows in dKV)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {  // single MUFU.EX2 (exp2f adds range fix-ups the softmax does not need)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5], degree-3 minimax polynomial for 2^f
// (max relative error 7.5e-5, far below the bf16 rounding the result gets), exponent patched in with one integer
// multiply-add. The attention kernels are bound by the 16 ex2 / clk / SM of the MUFU, so a fraction of the exponentials
// is evaluated this way (the FlashAttention-4 trick). x is clamped to >= -126 (masked scores are -inf).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;       // 1.5 * 2^23: round(x) lands in the low mantissa bits of t
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.055171649903059006f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 16 bf16 (8 columns of packed pairs), issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {  // 16-byte chunk in a [rows][128 B] SW128 tile
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void pair_sync(int quarter) {  // the two warps that share a TMEM lane quarter
  switch (quarter) {  // immediate barrier ids keep the CTA at 5 named barriers instead of reserving all 16
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

// Range metadata of one (plan, attention kind): built once per forward by egom2p_attn_ranges, shared by all layers,
// heads and by forward + backward. S = Mq rounded up to 128.
struct RangeMeta {
  int32_t* row_lo;      // (B, S) effective key range per query row; rows >= Mq: empty
  int32_t* row_hi;      // (B, S)
  float* row_scale;     // (B, S) scale*log2e, or 0 for fully-masked rows (uniform attention over all keys)
  int32_t* blk_lo;      // (B, S/64) union of the rows' ranges per 64-row block
  int32_t* blk_hi;
  int32_t* blk_lo_max;  // (B, S/64) intersection (INT_MAX / INT_MIN if the block holds a uniform or padding row)
  int32_t* blk_hi_min;
};
static inline int padS(int x) { return (x + 127) / 128 * 128; }  // row-metadata / lse stride: whole 128-row tiles
static inline int64_t align256(int64_t x) { return (x + 255) / 256 * 256; }
static inline int64_t range_meta_bytes(int B, int Mq) {
  const int64_t S = padS(Mq);
  return 3 * align256((int64_t)B * S * 4) + 4 * align256((int64_t)B * (S / 64) * 4);
}
static inline RangeMeta carve_meta(void* base, int B, int Mq) {
  const int64_t S = padS(Mq);
  const int64_t rows = align256((int64_t)B * S * 4), blks = align256((int64_t)B * (S / 64) * 4);
  char* p = reinterpret_cast<char*>(base);
  RangeMeta m;
  m.row_lo = reinterpret_cast<int32_t*>(p); p += rows;
  m.row_hi = reinterpret_cast<int32_t*>(p); p += rows;
  m.row_scale = reinterpret_cast<float*>(p); p += rows;
  m.blk_lo = reinterpret_cast<int32_t*>(p); p += blks;
  m.blk_hi = reinterpret_cast<int32_t*>(p); p += blks;
  m.blk_lo_max = reinterpret_cast<int32_t*>(p); p += blks;
  m.blk_hi_min = reinterpret_cast<int32_t*>(p);
  return m;
}

}  // namespace egom2p
