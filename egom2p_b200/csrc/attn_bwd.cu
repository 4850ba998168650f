// Backward of the range-masked flash attention (see attn.cu for the tiling and attn_common.cuh for the range metadata).
//   dQ  : per 128-query tile, loop over 64-key blocks: S, dP (TMEM) -> dS (smem) -> dQ += dS K (TMEM accumulator).
//         Also computes delta = rowsum(dO * O) for its rows and leaves -delta*scale in scratch for the dKV kernel.
//   dKV : per 128-key tile, loop over the 64-query blocks that can see it: S^T, dP^T (TMEM) -> P^T, dS^T (smem)
//         -> dV += P^T dO, dK += dS^T Q (TMEM accumulators). No atomics, deterministic.
// Both kernels use 8 math warps (two per TMEM lane quarter, 32 columns each) + TMA warp + MMA warp, 2 CTAs per SM.
#include "attn_common.cuh"

namespace egom2p {

// ------------------------------------------------------------------------------------------------ dQ
struct DqParams {
  int B, H, Mq, Nk, S;
  RangeMeta meta;
  const float* lse2;     // (B, H, S)
  const uint16_t* O;
  const uint16_t* dO;
  int64_t ldo;
  float* ndelta;         // (B, H, S) out: -delta * scale (natural-log units), consumed by the dKV kernel
  uint16_t* dQ;
  int64_t lddq;
};
constexpr int kDqStages = 3;
struct DqSmem {
  static constexpr int kQ = 0, kDO = kQ + kT * 128, kK = kDO + kT * 128, kV = kK + kDqStages * kBlk * 128,
                       kDS = kV + kDqStages * kBlk * 128, kX = kDS + kT * 128, kBar = kX + 2 * kT * 4,
                       kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
               const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const DqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sQ = smem + DqSmem::kQ, *sDO = smem + DqSmem::kDO, *sK = smem + DqSmem::kK, *sV = smem + DqSmem::kV,
          *sDS = smem + DqSmem::kDS;
  float* sX = reinterpret_cast<float*>(smem + DqSmem::kX);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DqSmem::kBar);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kDqStages;
  uint64_t* sdp_full = kv_empty + kDqStages;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* ds_full = sdp_free + 1;
  uint64_t* ds_empty = ds_full + 1;
  uint64_t* dq_full = ds_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dq_full + 1);
  int* s_range = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int trow = quarter * 32 + lane;
  const int row = q0 + trow;
  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kDqStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, kAttnComputeWarps);
    mbar_init(ds_full, kAttnComputeWarps);
    mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);
    fence_mbar_init();
    s_range[0] = INT_MAX;
    s_range[1] = INT_MIN;
  }
  if (warp == kMmaWarp) tmem_alloc<256>(tmem_slot);
  __syncthreads();
  int lo = INT_MAX, hi = INT_MIN;
  float rscale = 0.f;
  if (warp < kAttnComputeWarps && row < p.Mq) {
    lo = p.meta.row_lo[(int64_t)b * p.S + row];
    hi = p.meta.row_hi[(int64_t)b * p.S + row];
    rscale = p.meta.row_scale[(int64_t)b * p.S + row];
    if (half == 0 && hi > lo) { atomicMin(&s_range[0], lo); atomicMax(&s_range[1], hi); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lo_cta = s_range[0], hi_cta = s_range[1];
  const int nblk = hi_cta > lo_cta ? (hi_cta - lo_cta + kBlk - 1) / kBlk : 0;

  if (warp == kTmaWarp) {
    if (nblk > 0) {
      if (elect_one()) {
        mbar_expect_tx(q_full, 2 * kT * 128);
        tma_load_2d(sQ, &tmQ, q_full, h * kD, b * p.Mq + q0);
        tma_load_2d(sDO, &tmDO, q_full, h * kD, b * p.Mq + q0);
      }
      __syncwarp();
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kDqStages;
        mbar_wait(&kv_empty[st], ((j / kDqStages) & 1) ^ 1);
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
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBlk, 0, 0);
      constexpr uint32_t idesc_kmn = umma_idesc_bf16(128, kD, 0, 1);
      const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;
      const uint64_t dQ0 = umma_desc_kmajor_sw128(smem_u32(sQ)), dDO0 = umma_desc_kmajor_sw128(smem_u32(sDO));
      const uint64_t dDS0 = umma_desc_kmajor_sw128(smem_u32(sDS));
      auto issue_sdp = [&](int st) {
        const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sK + st * kBlk * 128));
        const uint64_t dV0 = umma_desc_kmajor_sw128(smem_u32(sV + st * kBlk * 128));
        if (elect_one()) {
          umma_bf16_ss(tS, dQ0, dK0, idesc_kk, 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tS, dQ0 + 2 * k, dK0 + 2 * k, idesc_kk, 1u);
          umma_bf16_ss(tDP, dDO0, dV0, idesc_kk, 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tDP, dDO0 + 2 * k, dV0 + 2 * k, idesc_kk, 1u);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_sdp(0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kDqStages;
        if (j + 1 < nblk) {  // next S / dP as soon as this block's values sit in registers
          const int st1 = (j + 1) % kDqStages;
          mbar_wait(sdp_free, j & 1);
          mbar_wait(&kv_full[st1], ((j + 1) / kDqStages) & 1);
          tc_fence_after();
          issue_sdp(st1);
        }
        mbar_wait(ds_full, j & 1);
        tc_fence_after();
        const uint64_t dKm0 = umma_desc_mnmajor_sw128(smem_u32(sK + st * kBlk * 128), 8192);
        if (elect_one()) {
          umma_bf16_ss(tDQ, dDS0, dKm0, idesc_kmn, j ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tDQ, dDS0 + 2 * k, dKm0 + 128 * k, idesc_kmn, 1u);
          umma_commit(ds_empty);
          umma_commit(&kv_empty[st]);
          if (j == nblk - 1) umma_commit(dq_full);
        }
        __syncwarp();
      }
    }
  } else {
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 32;
    const float rs_nat = rscale * kLn2;
    // delta = sum_d dO * O over the row (two halves combined through smem)
    float part = 0.f, nlse = -INFINITY;
    if (row < p.Mq) {
      const uint4* po = reinterpret_cast<const uint4*>(p.O + ((int64_t)b * p.Mq + row) * p.ldo + h * kD + half * 32);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dO + ((int64_t)b * p.Mq + row) * p.ldo + h * kD + half * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = po[i], g = pd[i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[u]));
          const float2 fg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[u]));
          part += fa.x * fg.x + fa.y * fg.y;
        }
      }
      nlse = -p.lse2[((int64_t)b * p.H + h) * p.S + row];
    }
    sX[half * kT + trow] = part;
    pair_sync(quarter);
    const float ndelta = -(part + sX[(half ^ 1) * kT + trow]) * rs_nat;
    if (half == 0 && row < p.S) p.ndelta[((int64_t)b * p.H + h) * p.S + row] = row < p.Mq ? ndelta : 0.f;

    for (int j = 0; j < nblk; ++j) {
      const int kv0 = lo_cta + j * kBlk + half * 32;
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t s[32], d[32];
      tmem_ld32(t_lane, s);
      tmem_ld32(t_lane + 64, d);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      float sc = rscale;
      if (!(rscale != 0.f && kv0 >= lo && kv0 + 32 <= hi)) {
        sc = rscale != 0.f ? rscale : 1.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const bool ok = (kv0 + c >= lo) && (kv0 + c < hi);
          s[c] = ok ? (rscale != 0.f ? s[c] : 0u) : 0xff800000u;
        }
      }
      if (j > 0) mbar_wait(ds_empty, (j - 1) & 1);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c8 * 8 + u;
          e[u] = ex2(fmaf(__uint_as_float(s[c]), sc, nlse)) * fmaf(__uint_as_float(d[c]), rs_nat, ndelta);
        }
        uint4 pk;
        pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
        *reinterpret_cast<uint4*>(sDS + swz_off(trow, half * 4 + c8)) = pk;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    uint32_t v[32];
    if (nblk > 0) {
      mbar_wait(dq_full, 0);
      tc_fence_after();
      tmem_ld32(t_lane + 128, v);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = 0u;
    }
    if (row < p.Mq) {
      uint16_t* out = p.dQ + ((int64_t)b * p.Mq + row) * p.lddq + h * kD + half * 32;
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
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
struct DkvParams {
  int B, H, Mq, Nk, S;
  float scale_log2;
  RangeMeta meta;
  const float* lse2;
  const float* ndelta;
  uint16_t* dK;
  uint16_t* dV;
  int64_t lddk, lddv;
};
constexpr int kDkvStages = 2;
constexpr int kMetaBytes = 5 * kBlk * 4;  // lse2, -delta*scale, lo, hi, row scale for 64 query rows
constexpr int kMaxQBlocks = 1024;
struct DkvSmem {
  static constexpr int kK = 0, kV = kK + kT * 128, kQ = kV + kT * 128, kDO = kQ + kDkvStages * kBlk * 128,
                       kPT = kDO + kDkvStages * kBlk * 128, kDST = kPT + kT * 128, kMeta = kDST + kT * 128,
                       kList = kMeta + kDkvStages * kMetaBytes, kBar = kList + kMaxQBlocks * 2, kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const DkvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sK = smem + DkvSmem::kK, *sV = smem + DkvSmem::kV, *sQ = smem + DkvSmem::kQ, *sDO = smem + DkvSmem::kDO,
          *sPT = smem + DkvSmem::kPT, *sDST = smem + DkvSmem::kDST, *sMeta = smem + DkvSmem::kMeta;
  uint16_t* s_list = reinterpret_cast<uint16_t*>(smem + DkvSmem::kList);  // bit 15: block is fully inside every row's range
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DkvSmem::kBar);
  uint64_t* kv_full = bars;
  uint64_t* q_full = bars + 1;
  uint64_t* q_empty = q_full + kDkvStages;
  uint64_t* sdp_full = q_empty + kDkvStages;
  uint64_t* sdp_free = sdp_full + 1;
  uint64_t* pds_full = sdp_free + 1;
  uint64_t* pds_empty = pds_full + 1;
  uint64_t* dkv_full = pds_empty + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dkv_full + 1);
  int* s_n = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int trow = quarter * 32 + lane;
  const int nqb = p.S / kBlk;
  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < kDkvStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, kAttnComputeWarps);
    mbar_init(pds_full, kAttnComputeWarps);
    mbar_init(pds_empty, 1);
    mbar_init(dkv_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) {  // query blocks whose key ranges intersect this key tile (ballot-compacted, ascending)
    int n = 0;
    for (int i0 = 0; i0 < nqb; i0 += 32) {
      const int i = i0 + lane;
      bool take = false, inside = false;
      if (i < nqb) {
        const int64_t o = (int64_t)b * nqb + i;
        take = p.meta.blk_hi[o] > kv0 && p.meta.blk_lo[o] < kv0 + kT;
        inside = p.meta.blk_lo_max[o] <= kv0 && p.meta.blk_hi_min[o] >= kv0 + kT;
      }
      const unsigned msk = __ballot_sync(0xffffffffu, take);
      const int pos = n + __popc(msk & ((1u << lane) - 1));
      if (take && pos < kMaxQBlocks) s_list[pos] = (uint16_t)(i | (inside ? 0x8000 : 0));
      n += __popc(msk);
    }
    if (lane == 0) *s_n = n < kMaxQBlocks ? n : kMaxQBlocks;
  }
  if (warp == kMmaWarp) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n = *s_n;

  if (warp == kTmaWarp) {
    if (n > 0) {
      if (elect_one()) {
        mbar_expect_tx(kv_full, 2 * kT * 128);
        tma_load_2d(sK, &tmK, kv_full, h * kD, b * p.Nk + kv0);
        tma_load_2d(sV, &tmV, kv_full, h * kD, b * p.Nk + kv0);
      }
      __syncwarp();
      for (int idx = 0; idx < n; ++idx) {
        const int st = idx % kDkvStages;
        const int r0 = (int)(s_list[idx] & 0x7fff) * kBlk;
        mbar_wait(&q_empty[st], ((idx / kDkvStages) & 1) ^ 1);
        uint8_t* meta = sMeta + st * kMetaBytes;
        const int64_t hoff = ((int64_t)b * p.H + h) * p.S + r0, roff = (int64_t)b * p.S + r0;
        if (elect_one()) {
          mbar_expect_tx(&q_full[st], 2 * kBlk * 128 + kMetaBytes);
          tma_load_2d(sQ + st * kBlk * 128, &tmQ, &q_full[st], h * kD, b * p.Mq + r0);
          tma_load_2d(sDO + st * kBlk * 128, &tmDO, &q_full[st], h * kD, b * p.Mq + r0);
          bulk_load(meta + 0 * 256, p.lse2 + hoff, 256, &q_full[st]);
          bulk_load(meta + 1 * 256, p.ndelta + hoff, 256, &q_full[st]);
          bulk_load(meta + 2 * 256, p.meta.row_lo + roff, 256, &q_full[st]);
          bulk_load(meta + 3 * 256, p.meta.row_hi + roff, 256, &q_full[st]);
          bulk_load(meta + 4 * 256, p.meta.row_scale + roff, 256, &q_full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == kMmaWarp) {
    if (n > 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBlk, 0, 0);
      constexpr uint32_t idesc_kmn = umma_idesc_bf16(128, kD, 0, 1);
      const uint32_t tST = tmem_base, tDPT = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 192;
      const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sK)), dV0 = umma_desc_kmajor_sw128(smem_u32(sV));
      const uint64_t dPT0 = umma_desc_kmajor_sw128(smem_u32(sPT)), dDST0 = umma_desc_kmajor_sw128(smem_u32(sDST));
      auto issue_sdp = [&](int st) {
        const uint64_t dQ0 = umma_desc_kmajor_sw128(smem_u32(sQ + st * kBlk * 128));
        const uint64_t dDO0 = umma_desc_kmajor_sw128(smem_u32(sDO + st * kBlk * 128));
        if (elect_one()) {
          umma_bf16_ss(tST, dK0, dQ0, idesc_kk, 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tST, dK0 + 2 * k, dQ0 + 2 * k, idesc_kk, 1u);
          umma_bf16_ss(tDPT, dV0, dDO0, idesc_kk, 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tDPT, dV0 + 2 * k, dDO0 + 2 * k, idesc_kk, 1u);
          umma_commit(sdp_full);
        }
        __syncwarp();
      };
      mbar_wait(kv_full, 0);
      mbar_wait(&q_full[0], 0);
      tc_fence_after();
      issue_sdp(0);
      for (int idx = 0; idx < n; ++idx) {
        const int st = idx % kDkvStages;
        if (idx + 1 < n) {
          const int st1 = (idx + 1) % kDkvStages;
          mbar_wait(sdp_free, idx & 1);
          mbar_wait(&q_full[st1], ((idx + 1) / kDkvStages) & 1);
          tc_fence_after();
          issue_sdp(st1);
        }
        mbar_wait(pds_full, idx & 1);
        tc_fence_after();
        const uint64_t dQm0 = umma_desc_mnmajor_sw128(smem_u32(sQ + st * kBlk * 128), 8192);
        const uint64_t dDOm0 = umma_desc_mnmajor_sw128(smem_u32(sDO + st * kBlk * 128), 8192);
        if (elect_one()) {
          umma_bf16_ss(tDV, dPT0, dDOm0, idesc_kmn, idx ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tDV, dPT0 + 2 * k, dDOm0 + 128 * k, idesc_kmn, 1u);
          umma_bf16_ss(tDK, dDST0, dQm0, idesc_kmn, idx ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < 4; ++k) umma_bf16_ss(tDK, dDST0 + 2 * k, dQm0 + 128 * k, idesc_kmn, 1u);
          umma_commit(pds_empty);
          umma_commit(&q_empty[st]);
          if (idx == n - 1) umma_commit(dkv_full);
        }
        __syncwarp();
      }
    }
  } else {
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 32;
    const int kidx = kv0 + trow;
    const float SC = p.scale_log2, RN = p.scale_log2 * kLn2;
    for (int idx = 0; idx < n; ++idx) {
      const int st = idx % kDkvStages;
      const bool inside = (s_list[idx] & 0x8000) != 0;
      mbar_wait(&q_full[st], (idx / kDkvStages) & 1);  // row metadata visible
      mbar_wait(sdp_full, idx & 1);
      tc_fence_after();
      const float* m_ls = reinterpret_cast<const float*>(sMeta + st * kMetaBytes) + half * 32;  // this thread's 32 query rows
      const float* m_nd = m_ls + 64;
      const int* m_lo = reinterpret_cast<const int*>(m_ls + 128);
      const int* m_hi = m_lo + 64;
      const float* m_rs = m_ls + 256;
      uint32_t s[32], d[32];
      tmem_ld32(t_lane, s);
      tmem_ld32(t_lane + 64, d);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sdp_free);
      if (idx > 0) mbar_wait(pds_empty, (idx - 1) & 1);
      if (inside) {
        // every (key, query) pair of this block is unmasked and every row uses the plain scale
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          const float4 l0 = *reinterpret_cast<const float4*>(m_ls + c8 * 8), l1 = *reinterpret_cast<const float4*>(m_ls + c8 * 8 + 4);
          const float4 n0 = *reinterpret_cast<const float4*>(m_nd + c8 * 8), n1 = *reinterpret_cast<const float4*>(m_nd + c8 * 8 + 4);
          const float ls[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
          const float nd[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
          float pt[8], e[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c8 * 8 + u;
            pt[u] = ex2(fmaf(__uint_as_float(s[c]), SC, -ls[u]));
            e[u] = pt[u] * fmaf(__uint_as_float(d[c]), RN, nd[u]);
          }
          uint4 pk;
          pk.x = pack_bf16(pt[0], pt[1]); pk.y = pack_bf16(pt[2], pt[3]); pk.z = pack_bf16(pt[4], pt[5]); pk.w = pack_bf16(pt[6], pt[7]);
          *reinterpret_cast<uint4*>(sPT + swz_off(trow, half * 4 + c8)) = pk;
          pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(sDST + swz_off(trow, half * 4 + c8)) = pk;
        }
      } else {
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float pt[8], e[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c8 * 8 + u;
            const float rs = m_rs[c];
            const bool ok = kidx >= m_lo[c] && kidx < m_hi[c];
            pt[u] = ok ? ex2(fmaf(rs != 0.f ? __uint_as_float(s[c]) : 0.f, rs, -m_ls[c])) : 0.f;
            e[u] = pt[u] * fmaf(__uint_as_float(d[c]), rs * kLn2, m_nd[c]);
          }
          uint4 pk;
          pk.x = pack_bf16(pt[0], pt[1]); pk.y = pack_bf16(pt[2], pt[3]); pk.z = pack_bf16(pt[4], pt[5]); pk.w = pack_bf16(pt[6], pt[7]);
          *reinterpret_cast<uint4*>(sPT + swz_off(trow, half * 4 + c8)) = pk;
          pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(sDST + swz_off(trow, half * 4 + c8)) = pk;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    if (n > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {  // 0: dV (cols 128..191), 1: dK (cols 192..255); this thread: 32 of the 64 columns
      uint32_t v[32];
      if (n > 0) {
        tmem_ld32(t_lane + 128 + which * 64, v);
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace egom2p

extern "C" int64_t egom2p_attn_bwd_scratch_bytes(int32_t B, int32_t H, int32_t Mq) {
  return (int64_t)B * H * egom2p::pad64(Mq) * 4 + 256;
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
  const int S = pad64(Mq);
  RangeMeta rm = carve_meta(const_cast<void*>(meta), B, Mq);
  float* ndelta = reinterpret_cast<float*>(scratch);

  CUtensorMap tmQ, tmDO, tmK, tmV;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kTotal);
    cudaError_t e2 = cudaFuncSetAttribute(attn_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kTotal);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_bwd: cudaFuncSetAttribute failed"); return EGOM2P_ERR_CUDA; }
    attr_set = true;
  }
  int rc;
  // ---- dQ (+ delta)
  if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kT, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kT, kD))) return rc;
  if (Nk > 0) {
    EGO_REQUIRE(K && V && dK && dV, "attn_bwd: K / V / dK / dV missing");
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kBlk, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kBlk, kD))) return rc;
  } else {
    tmK = tmQ; tmV = tmQ;
  }
  DqParams pq{B, H, Mq, Nk, S, rm, lse, O, dO, ldo, ndelta, dQ, lddq};
  attn_dq_kernel<<<dim3(S / kT + (S % kT ? 1 : 0), H, B), kAttnThreads, DqSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, pq);
  if ((rc = check_launch("attn_bwd dq"))) return rc;
  if (Nk == 0) return EGOM2P_OK;
  // ---- dK / dV
  EGO_REQUIRE(S / kBlk <= kMaxQBlocks, "attn_bwd: Mq too large (max %d)", kMaxQBlocks * kBlk);
  EGO_REQUIRE(lddk % 8 == 0 && lddv % 8 == 0 && ((uintptr_t)dK & 15) == 0 && ((uintptr_t)dV & 15) == 0, "attn_bwd: dK/dV alignment");
  if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kBlk, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kBlk, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kT, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kT, kD))) return rc;
  DkvParams pk{B, H, Mq, Nk, S, scale * kLog2e, rm, lse, ndelta, dK, dV, lddk, lddv};
  attn_dkv_kernel<<<dim3((Nk + kT - 1) / kT, H, B), kAttnThreads, DkvSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, pk);
  return check_launch("attn_bwd dkv");
}
