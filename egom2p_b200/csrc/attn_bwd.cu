// Backward of the range-masked flash attention (see attn.cu for the forward and attn_common.cuh for the metadata).
//
// One fused kernel: a CTA owns one tile of 128 keys of one (batch, head) and walks the 128-query blocks that can see
// it. Per block, five tcgen05 MMAs (all M = 128):
//     S  = Q K^T          (TMEM)        dP = dO V^T        (TMEM)
//     P  = exp2(S*c - lse), dS = P * (dP*scale - delta*scale)   (math warps: TMEM -> registers -> bf16 swizzled smem)
//     dV += P^T dO        (TMEM, A = P  read MN-major)
//     dK += dS^T Q        (TMEM, A = dS read MN-major)
//     dQ  = dS K          (TMEM, A = dS read K-major) -> drained to an fp32 accumulator in HBM by TMA reduce-add
// so S and dP are produced once per (key tile, query block) pair (10 tile-GEMM units instead of 14 for separate
// dQ / dKV kernels). Query rows sit on the TMEM lanes, so lse / delta / the key range are per-thread constants.
//
// 14 warps, 1 CTA / SM:
//   warps 0-7   math: warp w handles TMEM lane quarter w&3 (32 query rows) x key half (w>>2) (64 of the 128 columns)
//   warps 8-11  dQ drain: TMEM -> smem box -> cp.reduce.async.bulk.tensor (fp32 add), one 32-row quarter each
//   warp 12     TMA producer (K/V once, then the Q / dO / row-metadata ring);  warp 13: MMA issuer + TMEM owner
// A small pre-pass computes -delta*scale = -scale*rowsum(dO*O) and zeroes the dQ accumulator; a post-pass rounds the
// accumulator to bf16 into the caller's dQ.
#include "attn_common.cuh"

namespace egom2p {

constexpr int kBwdThreads = 448;
constexpr int kBwdStages = 3;
constexpr int kBwdMetaBytes = 5 * kT * 4;  // lse2, -delta*scale, lo, hi, row scale for 128 query rows
constexpr int kMaxQBlocks = 1024;

struct BwdParams {
  int B, H, Mq, Nk, S;
  RangeMeta meta;
  const float* lse2;    // (B, H, S)
  const float* ndelta;  // (B, H, S)  -delta * scale (natural-log units)
  uint16_t* dK;
  uint16_t* dV;
  int64_t lddk, lddv;
};
struct BwdSmem {
  static constexpr int kTile = kT * 128;  // 128 rows x 64 bf16
  static constexpr int kK = 0, kV = kK + kTile, kQ = kV + kTile, kDO = kQ + kBwdStages * kTile,
                       kP = kDO + kBwdStages * kTile,  // [128 q][2 blocks of 64 keys]: block stride kTile
                       kDS = kP + 2 * kTile, kBox = kDS + 2 * kTile,  // 4 drain warps x 4 KB (32 rows x 32 fp32)
                       kMeta = kBox + 4 * 4096, kList = kMeta + kBwdStages * kBwdMetaBytes,
                       kBar = kList + kMaxQBlocks * 2, kTotal = kBar + 256 + 1024;
};
static_assert(BwdSmem::kTotal <= 232448, "attention backward exceeds the 227 KB shared memory of one CTA");


// Applies the row's key range to 32 raw scores whose first key index is kv: masked -> -inf, uniform rows -> 0.
__device__ __forceinline__ void mask_scores(uint32_t (&s)[32], int kv, int lo, int hi, float rscale) {
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const bool ok = (kv + c >= lo) && (kv + c < hi);
    s[c] = ok ? (rscale != 0.f ? s[c] : 0u) : 0xff800000u;
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmDQ, const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sK = smem + BwdSmem::kK, *sV = smem + BwdSmem::kV, *sQ = smem + BwdSmem::kQ, *sDO = smem + BwdSmem::kDO,
          *sP = smem + BwdSmem::kP, *sDS = smem + BwdSmem::kDS, *sBox = smem + BwdSmem::kBox, *sMeta = smem + BwdSmem::kMeta;
  uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + BwdSmem::kList);  // bit 15: no masking needed for this block
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBar);
  uint64_t* kv_full = bars;                    // K / V tile landed
  uint64_t* q_full = bars + 1;                 // [3] Q / dO / metadata of a query block landed
  uint64_t* q_empty = q_full + kBwdStages;     // [3] dK MMA of the block retired (last reader of the stage)
  uint64_t* s_full = q_empty + kBwdStages;     // S in TMEM
  uint64_t* s_free = s_full + 1;               // math warps hold S in registers
  uint64_t* dp_full = s_free + 1;
  uint64_t* dp_free = dp_full + 1;
  uint64_t* p_full = dp_free + 1;              // P in smem
  uint64_t* p_empty = p_full + 1;              // dV MMA retired
  uint64_t* ds_full = p_empty + 1;             // dS in smem
  uint64_t* ds_empty = ds_full + 1;            // dK and dQ MMAs retired
  uint64_t* dq_full = ds_empty + 1;            // dQ block in TMEM
  uint64_t* dq_free = dq_full + 1;             // drain warps hold dQ in registers
  uint64_t* dkv_full = dq_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dkv_full + 1);
  int* s_n = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int nqb = p.S / kT;
  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < kBwdStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(s_full, 1);   mbar_init(s_free, kAttnComputeWarps);
    mbar_init(dp_full, 1);  mbar_init(dp_free, kAttnComputeWarps);
    mbar_init(p_full, kAttnComputeWarps);  mbar_init(p_empty, 1);
    mbar_init(ds_full, kAttnComputeWarps); mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);  mbar_init(dq_free, 4);
    mbar_init(dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {  // 128-query blocks whose key ranges intersect this key tile (ballot-compacted, ascending)
    int n = 0;
    for (int i0 = 0; i0 < nqb; i0 += 32) {
      const int i = i0 + lane;
      bool take = false, inside = false;
      if (i < nqb) {
        const int64_t o = ((int64_t)b * nqb + i) * 2;
        const int lo = min(p.meta.blk_lo[o], p.meta.blk_lo[o + 1]), hi = max(p.meta.blk_hi[o], p.meta.blk_hi[o + 1]);
        take = hi > kv0 && lo < kv0 + kT;
        inside = max(p.meta.blk_lo_max[o], p.meta.blk_lo_max[o + 1]) <= kv0 &&
                 min(p.meta.blk_hi_min[o], p.meta.blk_hi_min[o + 1]) >= kv0 + kT;
      }
      const unsigned msk = __ballot_sync(0xffffffffu, take);
      const int pos = n + __popc(msk & ((1u << lane) - 1));
      if (take && pos < kMaxQBlocks) s_list[pos] = (uint16_t)(i | (inside ? 0x8000 : 0));
      n += __popc(msk);
    }
    if (lane == 0) *s_n = n < kMaxQBlocks ? n : kMaxQBlocks;
  }
  if (warp == 13) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n = *s_n;
  constexpr uint32_t cS = 0, cDP = 128, cDV = 256, cDK = 320, cDQ = 384;  // TMEM columns

  if (warp >= 12) {
    if (warp == 12) {
      // ------------------------------------------------------------------------------------------ TMA producer
      if (n > 0) {
        if (elect_one()) {
          mbar_expect_tx(kv_full, 2 * BwdSmem::kTile);
          tma_load_2d(sK, &tmK, kv_full, h * kD, b * p.Nk + kv0);
          tma_load_2d(sV, &tmV, kv_full, h * kD, b * p.Nk + kv0);
        }
        __syncwarp();
        for (int idx = 0; idx < n; ++idx) {
          const int st = idx % kBwdStages;
          const int r0 = (int)(s_list[idx] & 0x7fff) * kT;
          mbar_wait(&q_empty[st], ((idx / kBwdStages) & 1) ^ 1);
          uint8_t* meta = sMeta + st * kBwdMetaBytes;
          const int64_t hoff = ((int64_t)b * p.H + h) * p.S + r0, roff = (int64_t)b * p.S + r0;
          if (elect_one()) {
            mbar_expect_tx(&q_full[st], 2 * BwdSmem::kTile + kBwdMetaBytes);
            tma_load_2d(sQ + st * BwdSmem::kTile, &tmQ, &q_full[st], h * kD, b * p.Mq + r0);
            tma_load_2d(sDO + st * BwdSmem::kTile, &tmDO, &q_full[st], h * kD, b * p.Mq + r0);
            bulk_load(meta + 0 * 512, p.lse2 + hoff, 512, &q_full[st]);
            bulk_load(meta + 1 * 512, p.ndelta + hoff, 512, &q_full[st]);
            bulk_load(meta + 2 * 512, p.meta.row_lo + roff, 512, &q_full[st]);
            bulk_load(meta + 3 * 512, p.meta.row_hi + roff, 512, &q_full[st]);
            bulk_load(meta + 4 * 512, p.meta.row_scale + roff, 512, &q_full[st]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 13) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      if (n > 0) {
        constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kT, 0, 0);    // S, dP: K-major x K-major, N = 128
        constexpr uint32_t idesc_mm = umma_idesc_bf16(128, kD, 1, 1);    // dV, dK: MN-major x MN-major, N = 64
        constexpr uint32_t idesc_km = umma_idesc_bf16(128, kD, 0, 1);    // dQ: K-major x MN-major, N = 64
        const uint32_t tS = tmem_base + cS, tDP = tmem_base + cDP, tDV = tmem_base + cDV, tDK = tmem_base + cDK,
                       tDQ = tmem_base + cDQ;
        const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sK)), dV0 = umma_desc_kmajor_sw128(smem_u32(sV));
        const uint64_t dKm0 = umma_desc_mnmajor_sw128(smem_u32(sK), 8192);
        const uint64_t dPm0 = umma_desc_mnmajor_sw128(smem_u32(sP), BwdSmem::kTile);
        const uint64_t dDSm0 = umma_desc_mnmajor_sw128(smem_u32(sDS), BwdSmem::kTile);
        const uint64_t dDSk0 = umma_desc_kmajor_sw128(smem_u32(sDS));
        const uint64_t dDSk1 = umma_desc_kmajor_sw128(smem_u32(sDS + BwdSmem::kTile));
        auto issue_s = [&](int st) {
          const uint64_t dQ0 = umma_desc_kmajor_sw128(smem_u32(sQ + st * BwdSmem::kTile));
          if (elect_one()) {
            umma_bf16_ss(tS, dQ0, dK0, idesc_kk, 0u);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_bf16_ss(tS, dQ0 + 2 * k, dK0 + 2 * k, idesc_kk, 1u);
            umma_commit(s_full);
          }
          __syncwarp();
        };
        auto issue_dp = [&](int st) {
          const uint64_t dDO0 = umma_desc_kmajor_sw128(smem_u32(sDO + st * BwdSmem::kTile));
          if (elect_one()) {
            umma_bf16_ss(tDP, dDO0, dV0, idesc_kk, 0u);
#pragma unroll
            for (int k = 1; k < 4; ++k) umma_bf16_ss(tDP, dDO0 + 2 * k, dV0 + 2 * k, idesc_kk, 1u);
            umma_commit(dp_full);
          }
          __syncwarp();
        };
        mbar_wait(kv_full, 0);
        mbar_wait(&q_full[0], 0);
        tc_fence_after();
        issue_s(0);
        issue_dp(0);
        for (int idx = 0; idx < n; ++idx) {
          const int st = idx % kBwdStages, st1 = (idx + 1) % kBwdStages;
          const uint32_t par = idx & 1;
          if (idx + 1 < n) {  // S of the next block as soon as this block's scores sit in registers
            mbar_wait(&q_full[st1], ((idx + 1) / kBwdStages) & 1);
            mbar_wait(s_free, par);
            tc_fence_after();
            issue_s(st1);
          }
          // dV += P^T dO
          mbar_wait(p_full, par);
          tc_fence_after();
          const uint64_t dDOm0 = umma_desc_mnmajor_sw128(smem_u32(sDO + st * BwdSmem::kTile), 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ss(tDV, dPm0 + 128 * k, dDOm0 + 128 * k, idesc_mm, (idx | k) ? 1u : 0u);
            umma_commit(p_empty);
          }
          __syncwarp();
          if (idx + 1 < n) {
            mbar_wait(dp_free, par);
            tc_fence_after();
            issue_dp(st1);
          }
          // dK += dS^T Q ; dQ = dS K
          mbar_wait(ds_full, par);
          tc_fence_after();
          const uint64_t dQm0 = umma_desc_mnmajor_sw128(smem_u32(sQ + st * BwdSmem::kTile), 8192);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k) umma_bf16_ss(tDK, dDSm0 + 128 * k, dQm0 + 128 * k, idesc_mm, (idx | k) ? 1u : 0u);
            umma_commit(&q_empty[st]);
          }
          __syncwarp();
          if (idx > 0) {
            mbar_wait(dq_free, (idx - 1) & 1);
            tc_fence_after();
          }
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16_ss(tDQ, (k < 4 ? dDSk0 : dDSk1) + 2 * (k & 3), dKm0 + 128 * k, idesc_km, k ? 1u : 0u);
            umma_commit(dq_full);
            umma_commit(ds_empty);
            if (idx == n - 1) umma_commit(dkv_full);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 8) {
    // -------------------------------------------------------------------------------------------- dQ drain
    const int quarter = warp & 3;
    uint8_t* box = sBox + quarter * 4096;
    const uint32_t t_dq = tmem_base + ((uint32_t)(quarter * 32) << 16) + cDQ;
    for (int idx = 0; idx < n; ++idx) {
      const int r0 = (int)(s_list[idx] & 0x7fff) * kT + quarter * 32;  // first query row (within the sample) of this box
      mbar_wait(dq_full, idx & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(t_dq, v0);
      tmem_ld32(t_dq + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (r0 >= p.Mq) continue;  // warp-uniform: rows past the sample's last query
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        if (lane == 0) tma_store_wait_read();  // the previous reduce has finished reading the box
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t* v = hb ? v1 : v0;
          *reinterpret_cast<uint4*>(box + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmDQ, box, h * kD + hb * 32, b * p.Mq + r0);
          tma_store_commit();
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  } else {
    // -------------------------------------------------------------------------------------------- math warps
    const int quarter = warp & 3, half = warp >> 2;
    const int trow = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int kcol0 = kv0 + half * 64;  // first key of this thread's 64 columns
    uint8_t* pP = sP + half * BwdSmem::kTile;
    uint8_t* pDS = sDS + half * BwdSmem::kTile;
    for (int idx = 0; idx < n; ++idx) {
      const int st = idx % kBwdStages;
      const uint32_t par = idx & 1;
      const uint16_t item = s_list[idx];
      const bool inside = (item & 0x8000) != 0;
      const int qrow = (int)(item & 0x7fff) * kT + trow;
      mbar_wait(&q_full[st], (idx / kBwdStages) & 1);  // row metadata visible
      const float* mf = reinterpret_cast<const float*>(sMeta + st * kBwdMetaBytes);
      const int* mi = reinterpret_cast<const int*>(mf);
      const bool row_ok = qrow < p.Mq;
      const float nlse = row_ok ? -mf[trow] : -INFINITY;
      const float ndl = mf[kT + trow];
      const int lo = mi[2 * kT + trow], hi = mi[3 * kT + trow];
      const float rscale = mf[4 * kT + trow];
      const float rs_nat = rscale * kLn2;
      // ---- P = exp2(S * c - lse), kept as packed bf16 pairs (the values the dV MMA consumes)
      uint32_t pp[32];
      {
        uint32_t s0[32], s1[32];
        mbar_wait(s_full, par);
        tc_fence_after();
        tmem_ld32(t_lane + cS + half * 64, s0);
        tmem_ld32(t_lane + cS + half * 64 + 32, s1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free);
        float sc = rscale;
        if (!inside) {
          if (!(rscale != 0.f && kcol0 >= lo && kcol0 + 64 <= hi)) {
            sc = rscale != 0.f ? rscale : 1.f;
            mask_scores(s0, kcol0, lo, hi, rscale);
            mask_scores(s1, kcol0 + 32, lo, hi, rscale);
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          pp[c] = pack_bf16(ex2(fmaf(__uint_as_float(s0[2 * c]), sc, nlse)), ex2(fmaf(__uint_as_float(s0[2 * c + 1]), sc, nlse)));
          pp[16 + c] = pack_bf16(ex2(fmaf(__uint_as_float(s1[2 * c]), sc, nlse)), ex2(fmaf(__uint_as_float(s1[2 * c + 1]), sc, nlse)));
        }
      }
      if (idx > 0) mbar_wait(p_empty, (idx - 1) & 1);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(pP + swz_off(trow, c8)) = make_uint4(pp[4 * c8], pp[4 * c8 + 1], pp[4 * c8 + 2], pp[4 * c8 + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- dS = P * (dP * scale - delta * scale), 32 columns at a time
      mbar_wait(dp_full, par);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t d[32], ds[16];
        tmem_ld32(t_lane + cDP + half * 64 + hh * 32, d);
        tmem_ld_wait();
        if (hh == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dp_free);
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const uint32_t w = pp[hh * 16 + c];
          const float p0 = __uint_as_float(w << 16), p1 = __uint_as_float(w & 0xffff0000u);
          ds[c] = pack_bf16(p0 * fmaf(__uint_as_float(d[2 * c]), rs_nat, ndl), p1 * fmaf(__uint_as_float(d[2 * c + 1]), rs_nat, ndl));
        }
        if (hh == 0 && idx > 0) mbar_wait(ds_empty, (idx - 1) & 1);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8)
          *reinterpret_cast<uint4*>(pDS + swz_off(trow, hh * 4 + c8)) = make_uint4(ds[4 * c8], ds[4 * c8 + 1], ds[4 * c8 + 2], ds[4 * c8 + 3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    // ---- epilogue: dV, dK (this thread: key row kv0 + trow, 32 of the 64 head-dim columns)
    if (n > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
    const int kidx = kv0 + trow;
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
      if (kidx < p.Nk) {
        uint16_t* out = (which == 0 ? p.dV + ((int64_t)b * p.Nk + kidx) * p.lddv : p.dK + ((int64_t)b * p.Nk + kidx) * p.lddk) +
                        h * kD + half * 32;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(v[c8 * 8 + 0]), __uint_as_float(v[c8 * 8 + 1]));
          pk.y = pack_bf16(__uint_as_float(v[c8 * 8 + 2]), __uint_as_float(v[c8 * 8 + 3]));
          pk.z = pack_bf16(__uint_as_float(v[c8 * 8 + 4]), __uint_as_float(v[c8 * 8 + 5]));
          pk.w = pack_bf16(__uint_as_float(v[c8 * 8 + 6]), __uint_as_float(v[c8 * 8 + 7]));
          reinterpret_cast<uint4*>(out)[c8] = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
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

extern "C" int64_t egom2p_attn_bwd_scratch_bytes(int32_t B, int32_t H, int32_t Mq) {
  using namespace egom2p;
  return align256((int64_t)B * H * padS(Mq) * 4) + align256((int64_t)B * Mq * H * kD * 4) + 256;
}

extern "C" int egom2p_attn_bwd(const uint16_t* Q, const uint16_t* K, const uint16_t* V, const uint16_t* O, const uint16_t* dO,
                               const float* lse, int32_t B, int32_t H, int32_t Mq, int32_t Nk, int64_t ldq, int64_t ldk,
                               int64_t ldv, int64_t ldo, const void* meta, float scale, void* scratch, uint16_t* dQ,
                               uint16_t* dK, uint16_t* dV, int64_t lddq, int64_t lddk, int64_t lddv, void* stream_) {
  using namespace egom2p;
  (void)scale;  // the per-row scale lives in the range metadata
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
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal);
      if (e != cudaSuccess) { set_error("attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return EGOM2P_ERR_CUDA; }
      attr_set = true;
    }
    CUtensorMap tmQ, tmDO, tmK, tmV, tmDQ;
    if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kT, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kT, kD))) return rc;
    if ((rc = make_tmap_2d(&tmDQ, dq_acc, 4, (uint64_t)B * Mq, (uint64_t)H * kD, (uint64_t)H * kD, 32, 32))) return rc;
    BwdParams pb{B, H, Mq, Nk, S, rm, lse, ndelta, dK, dV, lddk, lddv};
    attn_bwd_kernel<<<dim3((Nk + kT - 1) / kT, H, B), kBwdThreads, BwdSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, tmDQ, pb);
    if ((rc = check_launch("attn_bwd"))) return rc;
  }
  const int64_t rows = (int64_t)B * Mq;
  const int cols8 = H * kD / 8;
  const int64_t total = rows * cols8;
  attn_bwd_dq_cast_kernel<<<(unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, stream>>>(dq_acc, rows, cols8, dQ, lddq);
  return check_launch("attn_bwd dq cast");
}
