// Backward of the range-masked flash attention (see attn.cu for the forward and attn_common.cuh for the metadata).
//
// One fused kernel: a CTA owns one tile of 128 keys of one (batch, head) and walks the 128-query blocks that can see
// it. Keys sit on the TMEM lanes. Per block, five tcgen05 MMAs (all M = 128):
//     S^T  = K Q^T         (TMEM fp32)        dP^T = V dO^T       (TMEM fp32)
//     P^T  = exp2(S^T*c - lse_q), dS^T = P^T * (dP^T*scale - delta_q*scale)       (math warps)
//     (K and V sit in TMEM as packed bf16 -- copied once per CTA -- and are the TMEM A operands of the first two products)
//     dV  += P^T dO        A = P^T  read from TMEM (bf16, written with tcgen05.st in place over the thread's own S^T columns)
//     dK  += dS^T Q        A = dS^T read from TMEM (written in place over the dP^T columns)
//     dQ   = dS K          A = dS^T read MN-major from swizzled smem -> TMEM -> fp32 accumulator in HBM (TMA reduce-add)
// so S and dP are produced once per (key tile, query block) pair (10 tile-GEMM units instead of 14 for separate
// dQ / dKV kernels), and only dS touches shared memory: the kernel is bound by shared-memory bandwidth (MMA operand
// reads), so P^T / dS^T as TMEM operands matter more than instruction count.
//
// 14 warps, 1 CTA / SM:
//   warps 0-7   math: warp w handles TMEM lane quarter w&3 (32 keys) x query half w>>2 (64 of the 128 columns)
//   warps 8-11  dQ drain: TMEM -> smem box -> cp.reduce.async.bulk.tensor (fp32 add), one 32-row quarter each
//   warp 12     TMA producer (K/V once, then the Q / dO / row-metadata ring)
//   warp 13     MMA issuer A (S^T, dP^T, dQ) + TMEM owner;  warp 14: MMA issuer B (dV, dK)
// A small pre-pass computes -delta*scale = -scale*rowsum(dO*O) and zeroes the dQ accumulator; a post-pass rounds the
// accumulator to bf16 into the caller's dQ.
#include "attn_common.cuh"

namespace egom2p {

constexpr int kBwdMathWarps = 8;
constexpr int kBwdThreads = 32 * (kBwdMathWarps + 7);
constexpr int kDrainWarp0 = kBwdMathWarps, kBwdTmaWarp = kBwdMathWarps + 4, kBwdMmaWarp = kBwdMathWarps + 5,
              kBwdMmaWarpB = kBwdMathWarps + 6;
constexpr int kBwdStages = 3;
constexpr int kBwdMetaBytes = 5 * kT * 4;  // lse2, -delta*scale, lo, hi, row scale for 128 query rows
constexpr int kMaxQBlocks = 1024;

struct BwdParams {
  int B, H, Mq, Nk, S;
  RangeMeta meta;
  const float* lse2;    // (B, H, S)
  const float* ndelta;  // (B, H, S)  -delta * scale (natural-log units)
  float scale_log2;     // scale * log2(e) of normal rows
  uint16_t* dK;
  uint16_t* dV;
  int64_t lddk, lddv;
  int tma_out;  // dK / dV tiles leave through TMA stores (no rows past a sample's end: Nk % 128 == 0 or B == 1)
};
struct BwdSmem {
  static constexpr int kTile = kT * 128;  // 128 rows x 64 bf16
  static constexpr int kK = 0, kV = kK + kTile, kQ = kV + kTile, kDO = kQ + kBwdStages * kTile,
                       kDS = kDO + kBwdStages * kTile,  // dS^T: [128 keys][2 blocks of 64 queries], block stride kTile
                       kBox = kDS + 2 * kTile,          // 4 drain warps x 2 boxes x 4 KB (32 rows x 32 fp32)
                       kMeta = kBox + 8 * 4096, kList = kMeta + kBwdStages * kBwdMetaBytes,
                       kBar = kList + kMaxQBlocks * 2, kTotal = kBar + 256 + 1024;
};
static_assert(BwdSmem::kTotal <= 232448, "attention backward exceeds the 227 KB shared memory of one CTA");


#ifdef EGOM2P_TRACE
// Debug build only (python -m egom2p_b200.build with EGOM2P_TRACE=1): per-iteration clock stamps of CTA (0,0,0).
__device__ long long g_trace[16 * 64];
#define TRACE(slot, idx)                                                                          \
  do {                                                                                            \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == gridDim.z - 1 && (idx) < 64 && (threadIdx.x & 31) == 0) \
      g_trace[(slot) * 64 + (idx)] = clock64();                                                   \
  } while (0)
#else
#define TRACE(slot, idx) do {} while (0)
#endif

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float4 lds_f4(const void* p) {  // 16-byte shared load (a warp-wide broadcast when p is uniform)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmDQ, const __grid_constant__ CUtensorMap tmDK,
                const __grid_constant__ CUtensorMap tmDV, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sK = smem + BwdSmem::kK, *sV = smem + BwdSmem::kV, *sQ = smem + BwdSmem::kQ, *sDO = smem + BwdSmem::kDO,
          *sDS = smem + BwdSmem::kDS, *sBox = smem + BwdSmem::kBox, *sMeta = smem + BwdSmem::kMeta;
  uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + BwdSmem::kList);  // bit 15: block needs no masking; bit 14: every
                                                                          // row is a normal row (range masks only)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* kv_full = bars;                    // K / V tile landed
  uint64_t* q_full = bars + 1;                 // [3] Q / dO / metadata of a query block landed
  uint64_t* q_empty = q_full + kBwdStages;     // [3] dK MMA of the block retired (last reader of the stage)
  uint64_t* s_full = q_empty + kBwdStages;     // S in TMEM
  uint64_t* s_free = s_full + 1;               // math warps hold S in registers
  uint64_t* dp_full = s_free + 1;
  uint64_t* dp_free = dp_full + 1;             // (here: dK MMA retired -> the dP^T / dS^T columns may be overwritten)
  uint64_t* p_full = dp_free + 1;              // P in smem
  uint64_t* p_empty = p_full + 1;              // dV MMA retired
  uint64_t* ds_full = p_empty + 1;             // dS in smem
  uint64_t* ds_empty = ds_full + 1;            // dK and dQ MMAs retired
  uint64_t* dq_full = ds_empty + 1;            // dQ block in TMEM
  uint64_t* dq_free = dq_full + 1;             // drain warps hold dQ in registers
  uint64_t* dkv_full = dq_free + 1;
  uint64_t* kv_tmem = dkv_full + 1;             // K / V copied to TMEM by the math warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(kv_tmem + 1);
  int* s_n = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int nqb = p.S / kT;
  if (warp == 0) TRACE(5, 63);   // CTA start
  if (warp == kDrainWarp0 && lane == 0) {   // descriptor fetches (~1500 cycles from a cold TMA cache) off every first use
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
    tma_prefetch_desc(&tmDK);
    tma_prefetch_desc(&tmDV);
  }
  if (warp == kBwdTmaWarp && lane == 0) {
    // the K / V tile does not depend on the query-block list: its load is the first thing the CTA does and overlaps the
    // remaining barrier initialisation, the list build and the TMEM allocation
    mbar_init(kv_full, 1);
    fence_mbar_init();
    mbar_expect_tx(kv_full, 2 * BwdSmem::kTile);
    tma_load_2d(sK, &tmK, kv_full, h * kD, b * p.Nk + kv0);
    tma_load_2d(sV, &tmV, kv_full, h * kD, b * p.Nk + kv0);
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(s_full, 1);   mbar_init(s_free, kBwdMathWarps);
    mbar_init(dp_full, 1);  mbar_init(dp_free, 1);
    mbar_init(p_full, kBwdMathWarps);  mbar_init(p_empty, 1);
    mbar_init(ds_full, kBwdMathWarps); mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);  mbar_init(dq_free, 4);
    mbar_init(dkv_full, 1);
    mbar_init(kv_tmem, kBwdMathWarps);
    fence_mbar_init();
    TRACE(5, 57);   // K / V loads issued
  }
  // One query block's loads (Q, dO and the five per-row metadata vectors) into ring stage st.
  auto issue_q_block = [&](int qblk, int st) {
    const int r0 = qblk * kT;
    uint8_t* meta = sMeta + st * kBwdMetaBytes;
    const int64_t hoff = ((int64_t)b * p.H + h) * p.S + r0, roff = (int64_t)b * p.S + r0;
    mbar_expect_tx(&q_full[st], 2 * BwdSmem::kTile + kBwdMetaBytes);
    tma_load_2d(sQ + st * BwdSmem::kTile, &tmQ, &q_full[st], h * kD, b * p.Mq + r0);
    tma_load_2d(sDO + st * BwdSmem::kTile, &tmDO, &q_full[st], h * kD, b * p.Mq + r0);
    bulk_load(meta + 0 * 512, p.lse2 + hoff, 512, &q_full[st]);
    bulk_load(meta + 1 * 512, p.ndelta + hoff, 512, &q_full[st]);
    bulk_load(meta + 2 * 512, p.meta.row_lo + roff, 512, &q_full[st]);
    bulk_load(meta + 3 * 512, p.meta.row_hi + roff, 512, &q_full[st]);
    bulk_load(meta + 4 * 512, p.meta.row_scale + roff, 512, &q_full[st]);
  };
  if (warp == kBwdTmaWarp) {
    // The first query block of the list, found by this warp on its own (same test as the list build below), so that its
    // loads are in flight before the set-up barrier instead of one round trip after it.
    int first = -1;
    for (int i0 = 0; i0 < nqb && first < 0; i0 += 32) {
      const int i = i0 + lane;
      bool take = false;
      if (i < nqb) {
        const int64_t o = ((int64_t)b * nqb + i) * 2;
        const int lo = min(p.meta.blk_lo[o], p.meta.blk_lo[o + 1]), hi = max(p.meta.blk_hi[o], p.meta.blk_hi[o + 1]);
        take = hi > kv0 && lo < kv0 + kT;
      }
      const unsigned msk = __ballot_sync(0xffffffffu, take);
      if (msk) first = i0 + __ffs(msk) - 1;
    }
    __syncwarp();   // lane 0's barrier initialisation above is ordered before its own issue below
    if (first >= 0 && lane == 0) issue_q_block(first, 0);
  }
  if (warp == 0) {  // 128-query blocks whose key ranges intersect this key tile (ballot-compacted, ascending)
    int n = 0;
    for (int i0 = 0; i0 < nqb; i0 += 32) {
      const int i = i0 + lane;
      bool take = false, inside = false, normal = false;
      if (i < nqb) {
        const int64_t o = ((int64_t)b * nqb + i) * 2;
        const int lo = min(p.meta.blk_lo[o], p.meta.blk_lo[o + 1]), hi = max(p.meta.blk_hi[o], p.meta.blk_hi[o + 1]);
        take = hi > kv0 && lo < kv0 + kT;
        inside = max(p.meta.blk_lo_max[o], p.meta.blk_lo_max[o + 1]) <= kv0 &&
                 min(p.meta.blk_hi_min[o], p.meta.blk_hi_min[o + 1]) >= kv0 + kT;
        normal = p.meta.blk_lo_max[o] != INT_MAX && p.meta.blk_lo_max[o + 1] != INT_MAX;  // no uniform / padding row
      }
      const unsigned msk = __ballot_sync(0xffffffffu, take);
      const int pos = n + __popc(msk & ((1u << lane) - 1));
      if (take && pos < kMaxQBlocks) s_list[pos] = (uint16_t)(i | (inside ? 0x8000 : 0) | (normal ? 0x4000 : 0));
      n += __popc(msk);
    }
    if (lane == 0) *s_n = n < kMaxQBlocks ? n : kMaxQBlocks;
    TRACE(5, 55);   // list built
  }
  if (warp == kBwdMmaWarp) {
    tmem_alloc<512>(tmem_slot);
    TRACE(5, 54);   // TMEM allocated
  }
  if (warp == 1) TRACE(5, 53);   // an idle warp reaches the barrier
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n = *s_n;
  if (warp == 0) TRACE(5, 62);   // set-up barrier passed
  // TMEM columns. P^T (bf16 pairs) is written in place over the first 32 of each thread's 64 S^T columns (k-steps 0-3 at
  // columns 0-31, 4-7 at columns 64-95, like dS^T over dP^T), which frees 64 columns for K and V as packed bf16: the S^T and
  // dP^T products read their A operand from TMEM instead of shared memory (-32 KB of the 272 KB of shared-memory traffic per
  // 128 x 128 block that bound this kernel).
  constexpr uint32_t cS = 0, cDP = 128, cDV = 256, cDK = 320, cDQ = 384, cK = 448, cV = 480;

  if (warp >= kBwdTmaWarp) {
    if (warp == kBwdTmaWarp) {
      // ------------------------------------------------------------------------------------------ TMA producer
      if (n == 0) mbar_wait(kv_full, 0);   // nobody else consumes the early K / V load: it must land before the CTA exits
      if (n > 0) {
        for (int idx = 1; idx < n; ++idx) {   // block 0 was issued before the set-up barrier
          const int st = idx % kBwdStages;
          mbar_wait(&q_empty[st], ((idx / kBwdStages) & 1) ^ 1);
          if (elect_one()) issue_q_block((int)(s_list[idx] & 0x3fff), st);
          __syncwarp();
        }
      }
    } else if (warp == kBwdMmaWarp) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      if (n > 0) {
        constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kT, 0, 0);    // S^T, dP^T: K-major x K-major, N = 128
        constexpr uint32_t idesc_tm = umma_idesc_bf16(128, kD, 0, 1);    // dV, dK: A in TMEM x MN-major B, N = 64
        constexpr uint32_t idesc_mm = umma_idesc_bf16(128, kD, 1, 1);    // dQ: MN-major x MN-major, N = 64
        const uint32_t tS = tmem_base + cS, tDP = tmem_base + cDP, tDQ = tmem_base + cDQ, tK = tmem_base + cK, tV = tmem_base + cV;
        const uint64_t dKm0 = umma_desc_mnmajor_sw128(smem_u32(sK), 8192);
        const uint64_t dDSm0 = umma_desc_mnmajor_sw128(smem_u32(sDS), BwdSmem::kTile);
        auto issue_s = [&](int st) {
          const uint64_t dQ0 = umma_desc_kmajor_sw128(smem_u32(sQ + st * BwdSmem::kTile));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tS, tK + 8 * k, dQ0 + 2 * k, idesc_kk, k ? 1u : 0u);
            umma_commit(s_full);
          }
          __syncwarp();
        };
        auto issue_dp = [&](int st) {
          const uint64_t dDO0 = umma_desc_kmajor_sw128(smem_u32(sDO + st * BwdSmem::kTile));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tDP, tV + 8 * k, dDO0 + 2 * k, idesc_kk, k ? 1u : 0u);
            umma_commit(dp_full);
          }
          __syncwarp();
        };
        // One issuing thread sustains one tcgen05.mma per ~60 cycles whatever N <= 128 is (tools/micro/bench_umma.cu), so the
        // 32 MMAs of a block are split over two issuer warps: this one S^T, dP^T and dQ, warp B dV and dK. MMAs of
        // different issuers are not ordered against each other: every cross dependency goes through an mbarrier.
        mbar_wait(kv_tmem, 0);
        mbar_wait(&q_full[0], 0);
        tc_fence_after();
        issue_s(0);
        issue_dp(0);
        for (int idx = 0; idx < n; ++idx) {
          const int st1 = (idx + 1) % kBwdStages;
          const uint32_t par = idx & 1;
          if (idx + 1 < n) {  // S^T of the next block once this block's scores sit in registers AND its P^T (written over
                              // them) has been consumed by the dV product; the math warps are busy with dS^T meanwhile
            mbar_wait(&q_full[st1], ((idx + 1) / kBwdStages) & 1);
            mbar_wait(s_free, par);
            mbar_wait(p_empty, par);
            tc_fence_after();
            TRACE(0, idx);
            issue_s(st1);
          }
          // dQ = dS K (A = dS^T in smem, read MN-major)
          mbar_wait(ds_full, par);
          if (idx > 0) mbar_wait(dq_free, (idx - 1) & 1);
          tc_fence_after();
          TRACE(4, idx);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ss(tDQ, dDSm0 + 128 * k, dKm0 + 128 * k, idesc_mm, k ? 1u : 0u);
            umma_commit(dq_full);
            umma_commit(ds_empty);
          }
          __syncwarp();
          if (idx + 1 < n) {  // dP^T of the next block once warp B's dK MMA has consumed dS^T from the same columns
            mbar_wait(dp_free, par);
            tc_fence_after();
            TRACE(2, idx);
            issue_dp(st1);
          }
        }
      }
    } else if (warp == kBwdMmaWarpB) {
      // ------------------------------------------------------------------------------------------ MMA issuer B: dV, dK
      if (n > 0) {
        constexpr uint32_t idesc_tm = umma_idesc_bf16(128, kD, 0, 1);    // A in TMEM x MN-major B, N = 64
        const uint32_t tS = tmem_base + cS, tDP = tmem_base + cDP, tDV = tmem_base + cDV, tDK = tmem_base + cDK;
        for (int idx = 0; idx < n; ++idx) {
          const int st = idx % kBwdStages;
          const uint32_t par = idx & 1;
          // dV += P^T dO  (A: 8 k-steps of 16 queries = 8 TMEM columns each)
          mbar_wait(p_full, par);
          tc_fence_after();
          TRACE(1, idx);
          const uint64_t dDOm0 = umma_desc_mnmajor_sw128(smem_u32(sDO + st * BwdSmem::kTile), 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(tDV, tS + (k < 4 ? 8 * k : 64 + 8 * (k - 4)), dDOm0 + 128 * k, idesc_tm, (idx | k) ? 1u : 0u);
            umma_commit(p_empty);
          }
          __syncwarp();
          // dK += dS^T Q (A = dS^T, written in place over each thread's own dP^T columns: k-steps 0-3 at columns 0-31,
          // 4-7 at columns 64-95)
          mbar_wait(ds_full, par);
          tc_fence_after();
          TRACE(3, idx);
          const uint64_t dQm0 = umma_desc_mnmajor_sw128(smem_u32(sQ + st * BwdSmem::kTile), 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ts(tDK, tDP + (k < 4 ? 8 * k : 64 + 8 * (k - 4)), dQm0 + 128 * k, idesc_tm, (idx | k) ? 1u : 0u);
            umma_commit(&q_empty[st]);
            umma_commit(dp_free);
            if (idx == n - 1) umma_commit(dkv_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= kDrainWarp0) {
    // -------------------------------------------------------------------------------------------- dQ drain
    const int quarter = warp & 3;
    uint8_t* box0 = sBox + quarter * 8192;
    const uint32_t t_dq = tmem_base + ((uint32_t)(quarter * 32) << 16) + cDQ;
    for (int idx = 0; idx < n; ++idx) {
      const int r0 = (int)(s_list[idx] & 0x3fff) * kT + quarter * 32;  // first query row (within the sample) of this box
      mbar_wait(dq_full, idx & 1);
      tc_fence_after();
      if (warp == kDrainWarp0) TRACE(11, idx);
      const bool live = r0 < p.Mq;  // warp-uniform: rows past the sample's last query are skipped
      if (lane == 0) tma_store_wait_read();  // the previous block's reduces have finished reading both boxes
      __syncwarp();
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        uint32_t v[32];
        uint8_t* box = box0 + hb * 4096;
        tmem_ld32(t_dq + hb * 32, v);
        tmem_ld_wait();
        if (hb == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dq_free);
        }
        if (live) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmDQ, box, h * kD + hb * 32, b * p.Mq + r0);
            tma_store_commit();
          }
        }
      }
      if (warp == kDrainWarp0) TRACE(12, idx);
    }
    if (lane == 0) tma_store_wait_read();   // the boxes have left shared memory; the reduces complete on their own (kernel boundary)
  } else {
    // -------------------------------------------------------------------------------------------- math warps
    const int quarter = warp & 3, half = warp >> 2;  // half: which 64 of the block's 128 queries (TMEM columns)
    const int trow = quarter * 32 + lane;            // key row inside the tile == TMEM lane
    const int kidx = kv0 + trow;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint8_t* pDS = sDS + half * BwdSmem::kTile;
    const float SC = p.scale_log2, RN = p.scale_log2 * kLn2;
    if (n > 0) {   // K, V: smem -> TMEM (this thread's key row, its half of the head dim = 16 packed columns each)
      mbar_wait(kv_full, 0);
      if (warp == 0) TRACE(5, 56);   // K / V landed
      uint32_t kw[16], vw[16];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 a = *reinterpret_cast<const uint4*>(sK + swz_off(trow, half * 4 + c));
        const uint4 d = *reinterpret_cast<const uint4*>(sV + swz_off(trow, half * 4 + c));
        kw[4 * c] = a.x; kw[4 * c + 1] = a.y; kw[4 * c + 2] = a.z; kw[4 * c + 3] = a.w;
        vw[4 * c] = d.x; vw[4 * c + 1] = d.y; vw[4 * c + 2] = d.z; vw[4 * c + 3] = d.w;
      }
      tmem_st16(t_lane + cK + half * 16, kw);
      tmem_st16(t_lane + cV + half * 16, vw);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kv_tmem);
      if (warp == 0) TRACE(5, 58);   // K / V in TMEM
    }
    for (int idx = 0; idx < n; ++idx) {
      const int st = idx % kBwdStages;
      const uint32_t par = idx & 1;
      const bool inside = (s_list[idx] & 0x8000) != 0, normal = (s_list[idx] & 0x4000) != 0;
      // per-query metadata of this thread's 64 columns (warp-uniform addresses: broadcast loads)
      const float* m_ls = reinterpret_cast<const float*>(sMeta + st * kBwdMetaBytes) + half * 64;
      const float* m_nd = m_ls + kT;
      const int* m_lo = reinterpret_cast<const int*>(m_ls + 2 * kT);
      const int* m_hi = m_lo + kT;
      const float* m_rs = m_ls + 4 * kT;
      uint32_t pp[32];
      // ---- P^T = exp2(S^T * c - lse_q) as packed bf16 pairs
      {
        uint32_t s0[32], s1[32];
        if (warp == 0) TRACE(5, idx);
        mbar_wait(&q_full[st], (idx / kBwdStages) & 1);  // row metadata visible
        mbar_wait(s_full, par);
        tc_fence_after();
        if (warp == 0) TRACE(6, idx);
        tmem_ld32(t_lane + cS + half * 64, s0);
        tmem_ld32(t_lane + cS + half * 64 + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        if (inside) {
          // three passes (arguments, exponentials, packing) so that no instruction waits on the one just before it: with
          // two math warps per scheduler the MUFU latency is not hidden by other warps
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 la = lds_f4(m_ls + 4 * c4), lb = lds_f4(m_ls + 32 + 4 * c4);
            s0[4 * c4] = __float_as_uint(fmaf(__uint_as_float(s0[4 * c4]), SC, -la.x));
            s0[4 * c4 + 1] = __float_as_uint(fmaf(__uint_as_float(s0[4 * c4 + 1]), SC, -la.y));
            s0[4 * c4 + 2] = __float_as_uint(fmaf(__uint_as_float(s0[4 * c4 + 2]), SC, -la.z));
            s0[4 * c4 + 3] = __float_as_uint(fmaf(__uint_as_float(s0[4 * c4 + 3]), SC, -la.w));
            s1[4 * c4] = __float_as_uint(fmaf(__uint_as_float(s1[4 * c4]), SC, -lb.x));
            s1[4 * c4 + 1] = __float_as_uint(fmaf(__uint_as_float(s1[4 * c4 + 1]), SC, -lb.y));
            s1[4 * c4 + 2] = __float_as_uint(fmaf(__uint_as_float(s1[4 * c4 + 2]), SC, -lb.z));
            s1[4 * c4 + 3] = __float_as_uint(fmaf(__uint_as_float(s1[4 * c4 + 3]), SC, -lb.w));
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) {  // one exponential in four on the FMA pipes (ex2_poly), the rest on the MUFU
            s0[c] = __float_as_uint((c & 3) == 3 ? ex2_poly(__uint_as_float(s0[c])) : ex2(__uint_as_float(s0[c])));
            s1[c] = __float_as_uint((c & 3) == 3 ? ex2_poly(__uint_as_float(s1[c])) : ex2(__uint_as_float(s1[c])));
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            pp[c] = pack_bf16(__uint_as_float(s0[2 * c]), __uint_as_float(s0[2 * c + 1]));
            pp[16 + c] = pack_bf16(__uint_as_float(s1[2 * c]), __uint_as_float(s1[2 * c + 1]));
          }
        } else if (normal) {
          // range boundaries cross the block but every row is a normal row: same three passes, masked arguments -> -inf
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint32_t(&sv)[32] = g ? s1 : s0;
              const float4 l = lds_f4(m_ls + g * 32 + 4 * c4), lo = lds_f4(m_lo + g * 32 + 4 * c4), hi = lds_f4(m_hi + g * 32 + 4 * c4);
              const float lv[4] = {l.x, l.y, l.z, l.w};
              const int lov[4] = {__float_as_int(lo.x), __float_as_int(lo.y), __float_as_int(lo.z), __float_as_int(lo.w)};
              const int hiv[4] = {__float_as_int(hi.x), __float_as_int(hi.y), __float_as_int(hi.z), __float_as_int(hi.w)};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const bool ok = kidx >= lov[u] && kidx < hiv[u];
                sv[4 * c4 + u] = ok ? __float_as_uint(fmaf(__uint_as_float(sv[4 * c4 + u]), SC, -lv[u])) : 0xff800000u;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            s0[c] = __float_as_uint(ex2(__uint_as_float(s0[c])));
            s1[c] = __float_as_uint(ex2(__uint_as_float(s1[c])));
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            pp[c] = pack_bf16(__uint_as_float(s0[2 * c]), __uint_as_float(s0[2 * c + 1]));
            pp[16 + c] = pack_bf16(__uint_as_float(s1[2 * c]), __uint_as_float(s1[2 * c + 1]));
          }
        } else {  // block touches a range boundary, a uniform (fully masked) row or padding rows
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float e[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int q = 2 * c + u;
              const float rs = m_rs[q];
              const bool ok = kidx >= m_lo[q] && kidx < m_hi[q];
              const float sv = __uint_as_float(q < 32 ? s0[q & 31] : s1[q & 31]);
              e[u] = ok ? ex2(fmaf(rs != 0.f ? sv : 0.f, rs, -m_ls[q])) : 0.f;
            }
            pp[c] = pack_bf16(e[0], e[1]);
          }
        }
      }
      if (warp == 0) TRACE(7, idx);
      tmem_st32(t_lane + cS + half * 64, pp);    // in place over this thread's own (already loaded) S^T columns
      // dP^T of this block has been in TMEM for a while: start its load while the P^T store drains
      uint32_t d0[32];
      mbar_wait(dp_full, par);
      tc_fence_after();
      tmem_ld32(t_lane + cDP + half * 64, d0);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 0) TRACE(8, idx);
      // ---- dS^T = P^T * (dP^T * scale - delta_q * scale): the bracket in fp32, the product as one bf16x2 multiply per pair
      auto dsmul = [&](const uint32_t (&d)[32], int hh) {
        if (inside || normal) {  // every row uses the plain scale; straight-line: the eight metadata loads go back to back
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 nd = lds_f4(m_nd + hh * 32 + 4 * c4);
            const uint32_t t0 = pack_bf16(fmaf(__uint_as_float(d[4 * c4]), RN, nd.x), fmaf(__uint_as_float(d[4 * c4 + 1]), RN, nd.y));
            const uint32_t t1 = pack_bf16(fmaf(__uint_as_float(d[4 * c4 + 2]), RN, nd.z), fmaf(__uint_as_float(d[4 * c4 + 3]), RN, nd.w));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(pp[hh * 16 + 2 * c4]) : "r"(pp[hh * 16 + 2 * c4]), "r"(t0));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(pp[hh * 16 + 2 * c4 + 1]) : "r"(pp[hh * 16 + 2 * c4 + 1]), "r"(t1));
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 nd = lds_f4(m_nd + hh * 32 + 4 * c4);
            const float4 rs = lds_f4(m_rs + hh * 32 + 4 * c4);
            const uint32_t t0 = pack_bf16(fmaf(__uint_as_float(d[4 * c4]), rs.x * kLn2, nd.x), fmaf(__uint_as_float(d[4 * c4 + 1]), rs.y * kLn2, nd.y));
            const uint32_t t1 = pack_bf16(fmaf(__uint_as_float(d[4 * c4 + 2]), rs.z * kLn2, nd.z), fmaf(__uint_as_float(d[4 * c4 + 3]), rs.w * kLn2, nd.w));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(pp[hh * 16 + 2 * c4]) : "r"(pp[hh * 16 + 2 * c4]), "r"(t0));
            asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(pp[hh * 16 + 2 * c4 + 1]) : "r"(pp[hh * 16 + 2 * c4 + 1]), "r"(t1));
          }
        }
      };
      tmem_ld_wait();
      if (warp == 0) TRACE(9, idx);
      dsmul(d0, 0);
      if (warp == 0) TRACE(13, idx);
      tmem_ld32(t_lane + cDP + half * 64 + 32, d0);
      if (idx > 0) mbar_wait(ds_empty, (idx - 1) & 1);  // dQ MMA of the previous block retired: the smem copy of dS^T is free
      tmem_ld_wait();
      if (warp == 0) TRACE(14, idx);
      dsmul(d0, 1);
      if (warp == 0) TRACE(15, idx);
      // dS^T -> TMEM in place over this thread's own dP^T columns (all 64 are in registers), and -> smem for the dQ MMA
      tmem_st32(t_lane + cDP + half * 64, pp);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(pDS + swz_off(trow, c8)) = make_uint4(pp[4 * c8], pp[4 * c8 + 1], pp[4 * c8 + 2], pp[4 * c8 + 3]);
      fence_async_smem();
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (warp == 0) TRACE(10, idx);
    }
    // ---- epilogue: dV, dK (this thread: key row kv0 + trow, 32 of the 64 head-dim columns)
    if (n > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    if (warp == 0) TRACE(5, 61);   // dK / dV complete in TMEM
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      uint32_t v[32];
      if (n > 0) {
        tmem_ld32(t_lane + (which == 0 ? cDV : cDK) + half * 32, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0u;
      }
      uint4 pk[4];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        pk[c8].x = pack_bf16(__uint_as_float(v[c8 * 8 + 0]), __uint_as_float(v[c8 * 8 + 1]));
        pk[c8].y = pack_bf16(__uint_as_float(v[c8 * 8 + 2]), __uint_as_float(v[c8 * 8 + 3]));
        pk[c8].z = pack_bf16(__uint_as_float(v[c8 * 8 + 4]), __uint_as_float(v[c8 * 8 + 5]));
        pk[c8].w = pack_bf16(__uint_as_float(v[c8 * 8 + 6]), __uint_as_float(v[c8 * 8 + 7]));
      }
      if (p.tma_out) {
        // staged in the (idle) first Q / dO ring stages as swizzled 128 x 64 tiles and written by one TMA store each below:
        // 32 lanes storing 16 bytes into 32 different rows cost ~1000 cycles per CTA that nothing overlaps at one CTA per SM
        uint8_t* stage = which == 0 ? sQ : sDO;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(stage + swz_off(trow, half * 4 + c8)) = pk[c8];
        if (warp == 0) TRACE(5, 52 - which);
      } else if (kidx < p.Nk) {
        uint16_t* out = (which == 0 ? p.dV + ((int64_t)b * p.Nk + kidx) * p.lddv : p.dK + ((int64_t)b * p.Nk + kidx) * p.lddk) +
                        h * kD + half * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) reinterpret_cast<uint4*>(out)[c8] = pk[c8];
      }
    }
    if (p.tma_out) {
      fence_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight math warps
      if (warp == 0) TRACE(5, 50);
      if (warp == 0 && lane == 0) {
        tma_store_2d(&tmDV, sQ, h * kD, b * p.Nk + kv0);
        tma_store_2d(&tmDK, sDO, h * kD, b * p.Nk + kv0);
        tma_store_commit();
        tma_store_wait_read();   // the tiles have left shared memory; the writes complete on their own
      }
    }
    if (warp == 0) TRACE(5, 60);   // dK / dV written
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) TRACE(5, 59);     // CTA end
  if (warp == kBwdMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ pre / post passes
// One warp per query row (all heads): ndelta[b, h, r] = -scale_row * sum_d dO * O, and the row of the fp32 dQ
// accumulator is zeroed. Rows in [Mq, S) get ndelta = 0.
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const uint16_t* __restrict__ O, const uint16_t* __restrict__ dO,
                                                            int64_t ldo, const float* __restrict__ row_scale, int B, int H,
                                                            int Mq, int S, float* __restrict__ ndelta, float* __restrict__ dq_acc) {
  const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (int64_t)B * S) return;
  const int b = (int)(wid / S), r = (int)(wid % S);
  if (r >= Mq) {
    for (int hh = lane; hh < H; hh += 32) ndelta[((int64_t)b * H + hh) * S + r] = 0.f;
    return;
  }
  const float rs = row_scale[wid] * kLn2;
  const uint16_t* po = O + ((int64_t)b * Mq + r) * ldo;
  const uint16_t* pd = dO + ((int64_t)b * Mq + r) * ldo;
  float4* acc = reinterpret_cast<float4*>(dq_acc + ((int64_t)b * Mq + r) * (int64_t)H * kD);
  const int chunks = H * 8;  // 16-byte chunks of 8 bf16; 8 chunks per head
  for (int c0 = 0; c0 < chunks; c0 += 32) {
    const int c = c0 + lane;
    float part = 0.f;
    if (c < chunks) {
      const uint4 a = reinterpret_cast<const uint4*>(po)[c], g = reinterpret_cast<const uint4*>(pd)[c];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[u]));
        const float2 fg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[u]));
        part += fa.x * fg.x + fa.y * fg.y;
      }
      acc[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c < chunks && (lane & 7) == 0) ndelta[((int64_t)b * H + (c >> 3)) * S + r] = -part * rs;
  }
}

__global__ void __launch_bounds__(256) attn_bwd_dq_cast_kernel(const float* __restrict__ acc, int64_t rows, int cols8,
                                                               uint16_t* __restrict__ dq, int64_t lddq) {
  const int64_t total = rows * cols8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8;
    const int c = (int)(i % cols8);
    const float4 a = reinterpret_cast<const float4*>(acc + r * cols8 * 8)[2 * c];
    const float4 g = reinterpret_cast<const float4*>(acc + r * cols8 * 8)[2 * c + 1];
    uint4 pk;
    pk.x = pack_bf16(a.x, a.y); pk.y = pack_bf16(a.z, a.w); pk.z = pack_bf16(g.x, g.y); pk.w = pack_bf16(g.z, g.w);
    reinterpret_cast<uint4*>(dq + r * lddq)[c] = pk;
  }
}

}  // namespace egom2p

#ifdef EGOM2P_TRACE
extern "C" int egom2p_debug_attn_trace(long long* host_dst) {
  return (int)cudaMemcpyFromSymbol(host_dst, egom2p::g_trace, sizeof(egom2p::g_trace));
}
#endif

extern "C" int64_t egom2p_attn_bwd_scratch_bytes(int32_t B, int32_t H, int32_t Mq) {
  using namespace egom2p;
  return align256((int64_t)B * H * padS(Mq) * 4) + align256((int64_t)B * Mq * H * kD * 4) + 256;
}

extern "C" int egom2p_attn_bwd(const uint16_t* Q, const uint16_t* K, const uint16_t* V, const uint16_t* O, const uint16_t* dO,
                               const float* lse, int32_t B, int32_t H, int32_t Mq, int32_t Nk, int64_t ldq, int64_t ldk,
                               int64_t ldv, int64_t ldo, const void* meta, float scale, void* scratch, uint16_t* dQ,
                               uint16_t* dK, uint16_t* dV, int64_t lddq, int64_t lddk, int64_t lddv, void* stream_) {
  using namespace egom2p;
  cudaStream_t stream = (cudaStream_t)stream_;
  EGO_REQUIRE(Q && O && dO && lse && meta && scratch && dQ && B > 0 && H > 0 && Mq > 0 && Nk >= 0, "attn_bwd: bad argument");
  EGO_REQUIRE(((uintptr_t)scratch & 255) == 0 && ((uintptr_t)lse & 255) == 0 && ((uintptr_t)meta & 255) == 0,
              "attn_bwd: scratch / lse / meta must be 256-byte aligned");
  EGO_REQUIRE(lddq % 8 == 0 && ((uintptr_t)dQ & 15) == 0 && ldo % 8 == 0 && ((uintptr_t)O & 15) == 0 && ((uintptr_t)dO & 15) == 0,
              "attn_bwd: dQ / O / dO alignment");
  const int S = padS(Mq);
  EGO_REQUIRE(S / kT <= kMaxQBlocks, "attn_bwd: Mq too large (max %d)", kMaxQBlocks * kT);
  RangeMeta rm = carve_meta(const_cast<void*>(meta), B, Mq);
  float* ndelta = reinterpret_cast<float*>(scratch);
  float* dq_acc = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + align256((int64_t)B * H * S * 4));

  attn_bwd_prep_kernel<<<(unsigned)(((int64_t)B * S + 7) / 8), 256, 0, stream>>>(O, dO, ldo, rm.row_scale, B, H, Mq, S, ndelta, dq_acc);
  int rc = check_launch("attn_bwd prep");
  if (rc) return rc;
  if (Nk > 0) {
    EGO_REQUIRE(K && V && dK && dV, "attn_bwd: K / V / dK / dV missing");
    EGO_REQUIRE(lddk % 8 == 0 && lddv % 8 == 0 && ((uintptr_t)dK & 15) == 0 && ((uintptr_t)dV & 15) == 0, "attn_bwd: dK/dV alignment");
    static std::atomic<uint64_t> attr_done{0};
    if ((rc = ensure_dyn_smem(attn_bwd_kernel, BwdSmem::kTotal, attr_done, "attn_bwd"))) return rc;
    CUtensorMap tmQ, tmDO, tmK, tmV, tmDQ, tmDK, tmDV;
    if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kT, kD))) return rc;
    if ((rc = make_tmap_2d(&tmDQ, dq_acc, 4, (uint64_t)B * Mq, (uint64_t)H * kD, (uint64_t)H * kD, 32, 32))) return rc;
    const int tma_out = (Nk % kT == 0 || B == 1) ? 1 : 0;
    if (tma_out) {
      if ((rc = make_tmap_bf16_2d(&tmDK, dK, (uint64_t)B * Nk, (uint64_t)H * kD, lddk, kT, kD))) return rc;
      if ((rc = make_tmap_bf16_2d(&tmDV, dV, (uint64_t)B * Nk, (uint64_t)H * kD, lddv, kT, kD))) return rc;
    } else {
      tmDK = tmK;
      tmDV = tmV;
    }
    BwdParams pb{B, H, Mq, Nk, S, rm, lse, ndelta, scale * kLog2e, dK, dV, lddk, lddv, tma_out};
    attn_bwd_kernel<<<dim3((Nk + kT - 1) / kT, H, B), kBwdThreads, BwdSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, tmDQ, tmDK, tmDV, pb);
    if ((rc = check_launch("attn_bwd"))) return rc;
  }
  const int64_t rows = (int64_t)B * Mq;
  const int cols8 = H * kD / 8;
  const int64_t total = rows * cols8;
  attn_bwd_dq_cast_kernel<<<(unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, stream>>>(dq_acc, rows, cols8, dQ, lddq);
  return check_launch("attn_bwd dq cast");
}
