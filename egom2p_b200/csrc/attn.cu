// Range-masked flash attention for head_dim 64 on tcgen05/TMEM, operands staged by TMA (north-star kernel 2).
//
// Every mask the path produces is a contiguous key range per query row: encoder self / decoder
// cross = [0, n_enc[b]); decoder self = the row's own modality segment. Rows with an empty range reproduce the
// reference's masked_fill(-finfo.max) semantics (uniform attention over all Nk keys, A4/A5). The ranges are turned
// into per-row / per-block metadata once per forward (egom2p_attn_ranges) and shared by all layers and heads.
//
// One CTA = 128 query rows of one (batch, head); 2 CTAs are co-resident per SM so one CTA's softmax overlaps the
// other's MMAs. Warps 0-7: softmax math -- two warps per TMEM lane quarter, each owning 32 of the 64 score columns and
// 32 of the 64 output columns of its rows (row maxima are exchanged through smem); warp 8: TMA producer; warp 9:
// tcgen05.mma issuer + TMEM owner.
//
//   S = Q K^T (TMEM) -> online softmax -> P (bf16, swizzled smem) -> O_blk = P V (TMEM) -> O += in registers
//
// Two softmax paths, chosen per CTA:
//   * bound path (default): every row uses the FIXED reference m_r = |q_r| * max_k |k| * scale * log2(e) >= every score of the
//     row (Cauchy-Schwarz; max_k |k|^2 per (batch, head) comes from a small pre-pass over K). P = 2^(s - m_r) <= 1 can never
//     overflow, so there is no running maximum, no exchange between the two column halves of a row, no rescaling of O, and
//     the row sum comes out of the PV MMA itself (a constant tile of ones widens V to N = 80: column 64 of O is sum_k P).
//     The result is mathematically identical; P is merely scaled by 2^(max_r - m_r) >= 2^(-2 m_r) per row, which bf16 / fp32
//     represent exactly as well as long as m_r <= 60 (P stays a normal number). Per score that leaves one FFMA, one
//     MUFU.EX2, half a pack and 1/8 of a shared store: the kernel runs at the MUFU rate instead of the latency of the
//     max-exchange chain.
//   * online path (any row of the tile with m_r > 60, or no pre-pass scratch given): the running-maximum softmax with lazy
//     rescaling described above.
#include "attn_common.cuh"

namespace egom2p {

// ------------------------------------------------------------------------------------------------ range metadata
__global__ void __launch_bounds__(256) attn_rows_kernel(const int32_t* __restrict__ key_lo, const int32_t* __restrict__ key_hi,
                                                        int B, int Mq, int Nk, int S, float scale_log2, int empty_zero,
                                                        RangeMeta m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * S) return;
  const int b = (int)(i / S), r = (int)(i % S);
  int lo = INT_MAX / 2, hi = 0;
  float rs = 0.f;
  if (r < Mq) {
    lo = key_lo ? key_lo[(int64_t)b * Mq + r] : 0;
    hi = key_hi ? key_hi[(int64_t)b * Mq + r] : Nk;
    lo = max(lo, 0);
    hi = min(hi, Nk);
    rs = scale_log2;
    if (hi <= lo) {
      if (empty_zero) { lo = 0; hi = 0; }        // no key at all (the sampler's empty context): output exactly 0
      else { lo = 0; hi = Nk; rs = 0.f; }        // every key masked -> uniform over all keys (masked_fill(-finfo.max), A4)
    }
  }
  m.row_lo[i] = lo;
  m.row_hi[i] = hi;
  m.row_scale[i] = rs;
}
__global__ void __launch_bounds__(128) attn_blocks_kernel(int B, int S, RangeMeta m) {  // one warp per 64-row block
  const int blk = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (blk >= B * (S / 64)) return;
  int lo = INT_MAX, hi = INT_MIN, lo_max = INT_MIN, hi_min = INT_MAX;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int64_t r = (int64_t)blk * 64 + lane + 32 * k;
    const int l = m.row_lo[r], h = m.row_hi[r];
    if (h > l) { lo = min(lo, l); hi = max(hi, h); }
    const bool normal = (h > l) && m.row_scale[r] != 0.f;
    lo_max = normal ? max(lo_max, l) : INT_MAX;
    hi_min = normal ? min(hi_min, h) : INT_MIN;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    lo_max = max(lo_max, __shfl_xor_sync(0xffffffffu, lo_max, o));
    hi_min = min(hi_min, __shfl_xor_sync(0xffffffffu, hi_min, o));
  }
  if (lane == 0) { m.blk_lo[blk] = lo; m.blk_hi[blk] = hi; m.blk_lo_max[blk] = lo_max; m.blk_hi_min[blk] = hi_min; }
}

// max over the keys of one (batch, head) of |k|^2, as the int bits of a non-negative float (atomicMax), for the bound path.
// One warp handles 4 keys per step (8 lanes x 16 bytes = one 64-element key row).
__global__ void __launch_bounds__(256) attn_kmax_kernel(const uint16_t* __restrict__ K, int64_t ldk, int Nk, int H,
                                                        int* __restrict__ kmax2) {
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 256 + warp * 32;
  float best = 0.f;
  uint4 w[8];   // all eight loads in flight before the first reduction
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int key = k0 + it * 4 + (lane >> 3);
    w[it] = make_uint4(0u, 0u, 0u, 0u);
    if (key < Nk) w[it] = *reinterpret_cast<const uint4*>(K + ((int64_t)b * Nk + key) * ldk + h * kD + (lane & 7) * 8);
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const uint32_t ww[4] = {w[it].x, w[it].y, w[it].z, w[it].w};
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[u]));
      s = fmaf(f.x, f.x, fmaf(f.y, f.y, s));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    best = fmaxf(best, s);
  }
  best = warp_max(best);
  if (lane == 0) atomicMax(kmax2 + b * H + h, __float_as_int(best));
}

#ifdef EGOM2P_TRACE
// Debug build only (EGOM2P_TRACE=1 python -m egom2p_b200.build; tools/trace_attn_fwd.py): clock stamps of one late CTA.
__device__ long long g_fwd_trace[32];
#define FTRACE(slot)                                                                                              \
  do {                                                                                                            \
    if (blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == gridDim.z - 1 && (threadIdx.x & 31) == 0) g_fwd_trace[slot] = clock64(); \
  } while (0)
#else
#define FTRACE(slot) do {} while (0)
#endif

// ------------------------------------------------------------------------------------------------ forward
struct AttnFwdParams {
  int B, H, Mq, Nk, S;
  RangeMeta meta;
  uint16_t* O;
  int64_t ldo;
  int tma_out;  // the output tile leaves through one TMA store (rows past the sample's end must not exist: Mq % 128 == 0 or B == 1)
  float* lse2;  // (B, H, S), log2 domain: m + log2(l)
  const float* kmax2;  // (B, H) max_k |k|^2 (bound path), or NULL (online path only)
};
constexpr float kBoundLimit = 60.f;  // log2 units: rows with |q| max|k| scale log2e above this take the online path

constexpr int kFwdStages = 4;
constexpr float kRescaleThreshold = 8.f;  // log2 units: O / l are rescaled only when the row max grows by > 2^8
struct FwdSmem {
  static constexpr int kQ = 0;                             // staging only: Q moves to TMEM before the first MMA
  static constexpr int kK = kQ + kT * 128;
  static constexpr int kV = kK + kFwdStages * kBlk * 128;
  static constexpr int kOnes = kV + kFwdStages * kBlk * 128;  // [64 keys][64 bf16] of 1.0: V's MN group 64..127 (16 columns used)
  static constexpr int kX = kOnes + kBlk * 128;            // exchange slots: [2 parities][2 halves][128 rows] floats
  static constexpr int kBar = kX + 2 * 2 * kT * 4;
  static constexpr int kTotal = kBar + 256 + 1024;
};
// TMEM columns of one CTA (256 allocated; two CTAs per SM share the 512): S fp32 | O fp32 (64 head-dim columns + 16 of
// the ones group, column 64 = row sum of P) | Q as packed bf16 pairs | two P buffers as packed bf16 pairs.
constexpr int kPvN = kD + 16;
constexpr uint32_t cS = 0, cO = 64, cQ = cO + kPvN, cP = cQ + 32;
static_assert(cP + 2 * 32 <= 256, "attention forward TMEM budget");

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
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {  // one fp32 column of this thread's lane
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}

// Shared-memory bandwidth, not the MUFU, bounded the first version of this kernel (A = Q re-read from smem for every S
// product, P written to and read back from smem: 82 KB of smem traffic per 64-key block against 128 B / clk). Q and P now
// live in TMEM as packed bf16 pairs and enter the MMAs as TMEM operands; shared memory only carries K, V and the ones tile.
// O accumulates in TMEM across the whole key loop (tcgen05.mma accumulate), so the math warps never wait for a PV product
// inside the loop.
__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + FwdSmem::kQ;
  uint8_t* sK = smem + FwdSmem::kK;
  uint8_t* sV = smem + FwdSmem::kV;
  uint8_t* sOnes = smem + FwdSmem::kOnes;
  float* sX = reinterpret_cast<float*>(smem + FwdSmem::kX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* q_full = bars;                       // Q tile in smem (TMA)
  uint64_t* q_tmem = bars + 1;                   // Q copied to TMEM by the math warps, ones tile written
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + kFwdStages;
  uint64_t* s_full = kv_empty + kFwdStages;
  uint64_t* s_free = s_full + 1;
  uint64_t* p_full = s_free + 1;   // [2]
  uint64_t* pv_done = p_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
  int* s_slow = reinterpret_cast<int*>(tmem_slot + 1);   // set when some row of the tile exceeds the bound limit

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int trow = quarter * 32 + lane;  // row inside the tile (== TMEM lane)
  const int row = q0 + trow;
  if (warp == 0) FTRACE(0);   // CTA start

  if (warp == kMmaWarp && lane == 0) {   // descriptor fetches (~1500 cycles from a cold TMA cache) off every first use
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmQ);
    mbar_init(q_full, 1);
    mbar_init(q_tmem, kAttnComputeWarps);
    for (int i = 0; i < kFwdStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, kAttnComputeWarps);
    for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], kAttnComputeWarps); mbar_init(&pv_done[i], 1); }
    *s_slow = p.kmax2 ? 0 : 1;
    fence_mbar_init();
    // the Q tile does not depend on the key range: its load overlaps the range metadata reads and the TMEM allocation
    mbar_expect_tx(q_full, kT * 128);
    tma_load_2d(sQ, &tmQ, q_full, h * kD, b * p.Mq + q0);
  }
  if (warp == kMmaWarp) tmem_alloc<256>(tmem_slot);
  if (warp < kAttnComputeWarps) {   // the ones tile (constant): 512 x 16 bytes over 256 threads
    const uint4 one4 = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = one4;
    reinterpret_cast<uint4*>(sOnes)[threadIdx.x + 256] = one4;
  }
  int lo = INT_MAX, hi = INT_MIN;
  float rscale = 0.f;
  if (warp < kAttnComputeWarps && row < p.Mq) {   // the row's range: in flight while the barriers / TMEM are set up
    lo = p.meta.row_lo[(int64_t)b * p.S + row];
    hi = p.meta.row_hi[(int64_t)b * p.S + row];
    rscale = p.meta.row_scale[(int64_t)b * p.S + row];
  }
  // loaded here, not where it is used: behind the Q barrier this load's latency (~700 cycles) sat on the path to the first S
  const float kmax2 = (p.kmax2 && warp < kAttnComputeWarps) ? p.kmax2[b * p.H + h] : 0.f;
  // key range of the whole tile = union of its two 64-row blocks (precomputed by attn_blocks_kernel over the same rows)
  const int64_t bo = (int64_t)b * (p.S / 64) + 2 * blockIdx.x;
  const int lo_cta = min(p.meta.blk_lo[bo], p.meta.blk_lo[bo + 1]), hi_cta = max(p.meta.blk_hi[bo], p.meta.blk_hi[bo + 1]);
  const int nblk = hi_cta > lo_cta ? (hi_cta - lo_cta + kBlk - 1) / kBlk : 0;
  if (warp == kTmaWarp && lane == 0 && nblk > 0) {   // first K / V block: in flight before the set-up barrier as well
    mbar_expect_tx(&kv_full[0], 2 * kBlk * 128);
    tma_load_2d(sK, &tmK, &kv_full[0], h * kD, b * p.Nk + lo_cta);
    tma_load_2d(sV, &tmV, &kv_full[0], h * kD, b * p.Nk + lo_cta);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 0) FTRACE(1);   // set-up barrier passed (TMEM allocated, metadata loaded)

  if (warp == kTmaWarp) {
    // warp-uniform control flow (operands stay in uniform registers); one elected lane issues the copies
    if (nblk == 0) mbar_wait(q_full, 0);   // nobody else consumes the early Q load: it must land before the CTA exits
    if (nblk > 0) {
      for (int j = 1; j < nblk; ++j) {
        const int st = j % kFwdStages;
        mbar_wait(&kv_empty[st], ((j / kFwdStages) & 1) ^ 1);
        const int krow = b * p.Nk + lo_cta + j * kBlk;
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], 2 * kBlk * 128);
          tma_load_2d(sK + st * kBlk * 128, &tmK, &kv_full[st], h * kD, krow);
          tma_load_2d(sV + st * kBlk * 128, &tmV, &kv_full[st], h * kD, krow);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    if (nblk > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kBlk, 0, 0);    // A = Q (TMEM) x K-major K block
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kPvN, 0, 1);   // A = P (TMEM) x MN-major [V | ones]
      const uint32_t tS = tmem_base + cS, tO = tmem_base + cO, tQ = tmem_base + cQ, tP = tmem_base + cP;
      auto issue_s = [&](int st) {  // S = Q K^T, 4 k-steps of 16 (8 TMEM columns of packed pairs each)
        const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sK + st * kBlk * 128));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tS, tQ + 8 * k, dK0 + 2 * k, idesc_s, k ? 1u : 0u);
          umma_commit(s_full);
        }
        __syncwarp();
      };
      mbar_wait(q_tmem, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kFwdStages, buf = j & 1;
        if (j + 1 < nblk) {  // S(j+1) as soon as S(j) sits in registers: overlaps the softmax of block j
          const int st1 = (j + 1) % kFwdStages;
          mbar_wait(s_free, j & 1);
          mbar_wait(&kv_full[st1], ((j + 1) / kFwdStages) & 1);
          tc_fence_after();
          issue_s(st1);
        }
        mbar_wait(&p_full[buf], (j >> 1) & 1);
        tc_fence_after();
        // V stage = MN group 0 (head-dim columns 0..63); the ones tile = MN group 1, LBO bytes further on (columns 64..79)
        const uint64_t dV0 = umma_desc_mnmajor_sw128(smem_u32(sV + st * kBlk * 128), (uint32_t)(sOnes - (sV + st * kBlk * 128)));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ts(tO, tP + buf * 32 + 8 * k, dV0 + 128 * k, idesc_pv, (j | k) ? 1u : 0u);
          umma_commit(&pv_done[buf]);
          umma_commit(&kv_empty[st]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps: thread == (query row, column half)
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t t_s = t_lane + cS + half * 32;
    const uint32_t t_o = t_lane + cO + half * 32;
    const uint32_t t_p = t_lane + cP + half * 16;   // this thread's 32 keys = 16 packed columns of a P buffer
    float m_used = -INFINITY, l = 0.f;
    // ---- Q: smem -> TMEM (this thread's row, its half of the head dim = 16 packed columns), row norm for the bound path
    float m_row = 0.f;
    if (nblk > 0) {
      mbar_wait(q_full, 0);
      if (warp == 0) FTRACE(2);   // Q landed in smem
      float q2 = 0.f;
      uint32_t qw[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 w = *reinterpret_cast<const uint4*>(sQ + swz_off(trow, c));
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[u]));
          q2 = fmaf(f.x, f.x, fmaf(f.y, f.y, q2));
          if ((c >> 2) == half) qw[(c & 3) * 4 + u] = ww[u];
        }
      }
      tmem_st16(t_lane + cQ + half * 16, qw);
      if (p.kmax2) {
        m_row = sqrtf(q2 * kmax2) * rscale;   // 0 for uniform (fully masked) and padding rows
        if (m_row > kBoundLimit) *s_slow = 1;
      }
      tmem_st_wait();
      tc_fence_before();
      fence_async_smem();                                    // the ones tile -> visible to the MMA's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(q_tmem);
    }
    asm volatile("bar.sync 5, 256;" ::: "memory");            // the eight math warps: s_slow is final
    const bool fast = *s_slow == 0;
    if (warp == 0) FTRACE(3);     // Q in TMEM, path decided
    if (fast) {
      for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        const int kv0 = lo_cta + j * kBlk + half * 32;
        mbar_wait(s_full, j & 1);   // also: PV(j-2) has retired (commit order of the single MMA thread): P buffer `buf` is free
        tc_fence_after();
        if (warp == 0 && j == 0) FTRACE(4);          // S(0) ready
        if (warp == 0 && j == 1) FTRACE(5);          // S(1) ready
        if (warp == 0 && j == nblk - 1) FTRACE(6);   // S(last) ready
        uint32_t v[32];
        tmem_ld32(t_s, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        const float nm = -m_row;
        uint32_t pk[16];
        const bool interior = rscale != 0.f && kv0 >= lo && kv0 + 32 <= hi;
        if (__all_sync(0xffffffffu, interior)) {
          // every score of the warp's 32 x 32 patch is inside its row's range: a in [-2 m_r, 0]. (Moving a share of the
          // exponentials to the FMA pipes with ex2_poly was measured and changed nothing: profiles/r02_attn_fwd_ncu.md.)
#pragma unroll
          for (int c2 = 0; c2 < 16; ++c2)
            pk[c2] = pack_bf16(ex2(fmaf(__uint_as_float(v[c2 * 2]), rscale, nm)), ex2(fmaf(__uint_as_float(v[c2 * 2 + 1]), rscale, nm)));
        } else {
          const float sc = rscale != 0.f ? rscale : 1.f;
#pragma unroll
          for (int c2 = 0; c2 < 16; ++c2) {
            float e[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int c = c2 * 2 + u;
              const bool ok = (kv0 + c >= lo) && (kv0 + c < hi);
              const float sv = rscale != 0.f ? __uint_as_float(v[c]) : 0.f;
              e[u] = ok ? ex2(fmaf(sv, sc, nm)) : 0.f;
            }
            pk[c2] = pack_bf16(e[0], e[1]);
          }
        }
        tmem_st16(t_p + buf * 32, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[buf]);
      }
      m_used = m_row;
    } else
    for (int j = 0; j < nblk; ++j) {
      const int buf = j & 1;
      const int kv0 = lo_cta + j * kBlk + half * 32;  // first key of this thread's 32 columns
      mbar_wait(s_full, j & 1);   // (also covers PV(j-2): P buffer `buf` is free)
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_s, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float sc = rscale;
      if (!(rscale != 0.f && kv0 >= lo && kv0 + 32 <= hi)) {  // block touches the range boundary (or uniform row)
        sc = rscale != 0.f ? rscale : 1.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const bool ok = (kv0 + c >= lo) && (kv0 + c < hi);
          v[c] = ok ? (rscale != 0.f ? v[c] : 0u) : 0xff800000u;
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY, bm2 = -INFINITY, bm3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        bm0 = fmaxf(bm0, __uint_as_float(v[c]));
        bm1 = fmaxf(bm1, __uint_as_float(v[c + 1]));
        bm2 = fmaxf(bm2, __uint_as_float(v[c + 2]));
        bm3 = fmaxf(bm3, __uint_as_float(v[c + 3]));
      }
      float bm = fmaxf(fmaxf(bm0, bm1), fmaxf(bm2, bm3)) * sc;  // sc > 0, so max commutes with the scaling
      float* xs = sX + (j & 1) * 2 * kT;
      xs[half * kT + trow] = bm;
      pair_sync(quarter);
      bm = fmaxf(bm, xs[(half ^ 1) * kT + trow]);
      // lazy rescale: both half-warps of a row see the same bm / m_used, so they take the same decision
      const bool need = bm > m_used + kRescaleThreshold;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2(m_used - bm) : 1.f;
        if (need) { m_used = bm; l *= alpha; }
        if (j > 0) {
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // PV(j-1) has landed: O is stable
          tc_fence_after();
          uint32_t ov[32];
          tmem_ld32(t_o, ov);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) ov[c] = __float_as_uint(__uint_as_float(ov[c]) * alpha);
          tmem_st32(t_o, ov);
          tmem_st_wait();
          tc_fence_before();
        }
      }
      const float nm = (m_used == -INFINITY) ? 0.f : -m_used;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
      uint32_t pk[16];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) e[u] = ex2(fmaf(__uint_as_float(v[c8 * 8 + u]), sc, nm));
        sum0 += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
        pk[c8 * 4 + 0] = pack_bf16(e[0], e[1]); pk[c8 * 4 + 1] = pack_bf16(e[2], e[3]);
        pk[c8 * 4 + 2] = pack_bf16(e[4], e[5]); pk[c8 * 4 + 3] = pack_bf16(e[6], e[7]);
      }
      l += (sum0 + sum1) + (sum2 + sum3);
      tmem_st16(t_p + buf * 32, pk);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[buf]);
    }
    // epilogue: O / l
    if (warp == 0) FTRACE(7);     // last P published
    uint32_t ov[32];
    if (nblk > 0) {
      mbar_wait(&pv_done[(nblk - 1) & 1], ((nblk - 1) >> 1) & 1);
      tc_fence_after();
      if (warp == 0) FTRACE(8);   // last PV retired
      tmem_ld32(t_o, ov);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) ov[c] = 0u;
    }
    float lt;
    if (fast) {
      lt = nblk > 0 ? __uint_as_float(tmem_ld1(t_lane + cO + kD)) : 0.f;
      tmem_ld_wait();
    } else {
      float* xs = sX + (nblk & 1) * 2 * kT;
      xs[half * kT + trow] = l;
      pair_sync(quarter);
      lt = l + xs[(half ^ 1) * kT + trow];
    }
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    uint4 pk4[4];
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      pk4[c8].x = pack_bf16(__uint_as_float(ov[c8 * 8 + 0]) * inv, __uint_as_float(ov[c8 * 8 + 1]) * inv);
      pk4[c8].y = pack_bf16(__uint_as_float(ov[c8 * 8 + 2]) * inv, __uint_as_float(ov[c8 * 8 + 3]) * inv);
      pk4[c8].z = pack_bf16(__uint_as_float(ov[c8 * 8 + 4]) * inv, __uint_as_float(ov[c8 * 8 + 5]) * inv);
      pk4[c8].w = pack_bf16(__uint_as_float(ov[c8 * 8 + 6]) * inv, __uint_as_float(ov[c8 * 8 + 7]) * inv);
    }
    if (p.tma_out) {
      // The 128 x 64 bf16 tile goes through the (now idle) Q staging buffer and leaves as ONE bulk tensor store: 32 lanes
      // writing 16 bytes each into 32 different rows cost ~1000 LSU cycles per CTA as direct global stores.
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) *reinterpret_cast<uint4*>(sQ + swz_off(trow, half * 4 + c8)) = pk4[c8];
      fence_async_smem();
      asm volatile("bar.sync 5, 256;" ::: "memory");
      if (warp == 0 && lane == 0) {
        tma_store_2d(&tmO, sQ, h * kD, b * p.Mq + q0);
        tma_store_commit();
        tma_store_wait_read();
      }
    } else if (row < p.Mq) {
      uint16_t* orow = p.O + ((int64_t)b * p.Mq + row) * p.ldo + h * kD + half * 32;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) reinterpret_cast<uint4*>(orow)[c8] = pk4[c8];
    }
    if (row < p.Mq && p.lse2 && half == 0) p.lse2[((int64_t)b * p.H + h) * p.S + row] = lt > 0.f ? m_used + log2f(lt) : INFINITY;
  }
  if (warp == 0) FTRACE(9);       // outputs stored
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
  if (warp == 0) FTRACE(10);      // CTA end
}

}  // namespace egom2p

#ifdef EGOM2P_TRACE
extern "C" int egom2p_debug_attn_fwd_trace(long long* host_dst) {
  return (int)cudaMemcpyFromSymbol(host_dst, egom2p::g_fwd_trace, sizeof(egom2p::g_fwd_trace));
}
#endif

extern "C" int egom2p_attn_lse_stride(int32_t Mq) { return egom2p::padS(Mq); }
extern "C" int64_t egom2p_attn_ranges_bytes(int32_t B, int32_t Mq) { return egom2p::range_meta_bytes(B, Mq); }

extern "C" int egom2p_attn_ranges(const int32_t* key_lo, const int32_t* key_hi, int32_t B, int32_t Mq, int32_t Nk, float scale,
                                  int32_t empty_zero, void* meta, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(meta && B > 0 && Mq > 0 && Nk >= 0, "attn_ranges: bad argument");
  EGO_REQUIRE((key_lo == nullptr) == (key_hi == nullptr), "attn_ranges: key_lo / key_hi must both be given or both NULL");
  EGO_REQUIRE(((uintptr_t)meta & 255) == 0, "attn_ranges: meta must be 256-byte aligned");
  const int S = padS(Mq);
  RangeMeta m = carve_meta(meta, B, Mq);
  attn_rows_kernel<<<(unsigned)(((int64_t)B * S + 255) / 256), 256, 0, (cudaStream_t)stream>>>(key_lo, key_hi, B, Mq, Nk, S,
                                                                                           scale * kLog2e, empty_zero, m);
  int rc = check_launch("attn_ranges rows");
  if (rc) return rc;
  attn_blocks_kernel<<<(B * (S / 64) + 3) / 4, 128, 0, (cudaStream_t)stream>>>(B, S, m);
  return check_launch("attn_ranges blocks");
}

extern "C" int egom2p_attn_fwd(const uint16_t* Q, const uint16_t* K, const uint16_t* V, int32_t B, int32_t H, int32_t Mq,
                               int32_t Nk, int64_t ldq, int64_t ldk, int64_t ldv, const void* meta, uint16_t* O, int64_t ldo,
                               float* lse, float* kmax_scratch, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(Q && O && meta && B > 0 && H > 0 && Mq > 0 && Nk >= 0, "attn_fwd: bad argument");
  EGO_REQUIRE(ldo % 8 == 0 && ((uintptr_t)O & 15) == 0, "attn_fwd: O must be 16-byte aligned with ldo %% 8 == 0");
  AttnFwdParams p{};
  p.B = B; p.H = H; p.Mq = Mq; p.Nk = Nk; p.S = padS(Mq);
  p.meta = carve_meta(const_cast<void*>(meta), B, Mq);
  p.O = O; p.ldo = ldo; p.lse2 = lse;
  p.kmax2 = nullptr;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kT, kD);
  if (rc) return rc;
  p.tma_out = ((Mq % kT == 0 || B == 1) && ((uintptr_t)O & 15) == 0 && (ldo * 2) % 16 == 0) ? 1 : 0;
  if (p.tma_out) {
    if ((rc = make_tmap_bf16_2d(&tmO, O, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kT, kD))) return rc;
  } else {
    tmO = tmQ;
  }
  if (Nk > 0) {
    EGO_REQUIRE(K && V, "attn_fwd: K / V missing");
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kBlk, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kBlk, kD))) return rc;
  } else {
    tmK = tmQ;
    tmV = tmQ;
  }
  if (kmax_scratch && Nk > 0) {   // pre-pass of the bound path: max_k |k|^2 per (batch, head)
    EGO_REQUIRE(ldk % 8 == 0 && ((uintptr_t)K & 15) == 0, "attn_fwd: K must be 16-byte aligned with ldk %% 8 == 0");
    cudaError_t e = cudaMemsetAsync(kmax_scratch, 0, (size_t)B * H * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("attn_fwd: memset: %s", cudaGetErrorString(e)); return EGOM2P_ERR_CUDA; }
    attn_kmax_kernel<<<dim3((Nk + 255) / 256, H, B), 256, 0, (cudaStream_t)stream>>>(K, ldk, Nk, H, reinterpret_cast<int*>(kmax_scratch));
    if ((rc = check_launch("attn_kmax"))) return rc;
    p.kmax2 = kmax_scratch;
  }
  dim3 grid((Mq + kT - 1) / kT, H, B);
  static std::atomic<uint64_t> attr_done{0};
  if ((rc = ensure_dyn_smem(attn_fwd_kernel, FwdSmem::kTotal, attr_done, "attn_fwd"))) return rc;
  attn_fwd_kernel<<<grid, kAttnThreads, FwdSmem::kTotal, (cudaStream_t)stream>>>(tmQ, tmK, tmV, tmO, p);
  return check_launch("attn_fwd");
}
