// Backward of the range-masked flash attention (see attn.cu for the tiling): three kernels.
//   prep : delta = rowsum(dO * O), per-row effective key ranges / scales, per-64-row-block range summaries
//   dQ   : per 128-query tile, loop over 64-key blocks: S, dP (TMEM) -> dS (smem) -> dQ += dS K (TMEM accumulator)
//   dKV  : per 128-key tile, loop over the 64-query blocks that can see it: S^T, dP^T (TMEM) -> P^T, dS^T (smem)
//          -> dV += P^T dO, dK += dS^T Q (TMEM accumulators). No atomics, deterministic.
#include <climits>

#include "common.cuh"

namespace egom2p {

constexpr int kThreads = 192;
constexpr int kD = 64;
constexpr int kT = 128;   // tile rows (queries in dQ, keys in dKV)
constexpr int kBlk = 64;  // inner block
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct Scratch {
  float* delta;      // (B, H, S)   -delta * rs_nat  (pre-folded for  dS = P * fma(dP, rs_nat, .))
  float* nlse;       // (B, H, S)   -lse2
  int32_t* blk_lo_max;  // (B, S/64)  max row lo  (INT_MAX if the block holds a uniform / padding row)
  int32_t* blk_hi_min;  // (B, S/64)  min row hi
  int32_t* row_lo;   // (B, S)
  int32_t* row_hi;   // (B, S)
  float* row_scale;  // (B, S)  scale*log2e, or 0 for fully-masked (uniform) rows
  int32_t* blk_lo;   // (B, S/64)
  int32_t* blk_hi;   // (B, S/64)
};
static inline int pad64(int x) { return (x + 63) / 64 * 64; }
static inline Scratch carve(void* base, int B, int H, int Mq) {
  const int64_t S = pad64(Mq);
  Scratch s;
  char* p = reinterpret_cast<char*>(base);
  s.delta = reinterpret_cast<float*>(p); p += (int64_t)B * H * S * 4;
  s.nlse = reinterpret_cast<float*>(p); p += (int64_t)B * H * S * 4;
  s.blk_lo_max = reinterpret_cast<int32_t*>(p); p += (int64_t)B * (S / 64) * 4 + 192;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
  s.blk_hi_min = reinterpret_cast<int32_t*>(p); p += (int64_t)B * (S / 64) * 4 + 192;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 255) & ~(uintptr_t)255);
  s.row_lo = reinterpret_cast<int32_t*>(p); p += (int64_t)B * S * 4;
  s.row_hi = reinterpret_cast<int32_t*>(p); p += (int64_t)B * S * 4;
  s.row_scale = reinterpret_cast<float*>(p); p += (int64_t)B * S * 4;
  s.blk_lo = reinterpret_cast<int32_t*>(p); p += (int64_t)B * (S / 64) * 4;
  s.blk_hi = reinterpret_cast<int32_t*>(p);
  return s;
}

// ------------------------------------------------------------------------------------------------ prep
__global__ void __launch_bounds__(256) attn_prep_kernel(const uint16_t* __restrict__ O, const uint16_t* __restrict__ dO,
                                                        int64_t ldo, int B, int H, int Mq, int Nk, int S,
                                                        const int32_t* __restrict__ key_lo, const int32_t* __restrict__ key_hi,
                                                        const float* __restrict__ lse2, float scale_log2, Scratch sc) {
  const int64_t gw = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);  // one warp per (b, padded row)
  if (gw >= (int64_t)B * S) return;
  const int lane = threadIdx.x & 31;
  const int b = (int)(gw / S), r = (int)(gw % S);
  if (r >= Mq) {
    if (lane == 0) { sc.row_lo[gw] = INT_MAX / 2; sc.row_hi[gw] = 0; sc.row_scale[gw] = 0.f; }
    for (int h = lane; h < H; h += 32) {
      sc.delta[((int64_t)b * H + h) * S + r] = 0.f;
      sc.nlse[((int64_t)b * H + h) * S + r] = -INFINITY;
    }
    return;
  }
  int lo = key_lo ? key_lo[(int64_t)b * Mq + r] : 0;
  int hi = key_hi ? key_hi[(int64_t)b * Mq + r] : Nk;
  lo = max(lo, 0); hi = min(hi, Nk);
  float rs = scale_log2;
  if (hi <= lo) { lo = 0; hi = Nk; rs = 0.f; }
  if (lane == 0) { sc.row_lo[gw] = lo; sc.row_hi[gw] = hi; sc.row_scale[gw] = rs; }
  const uint32_t* o = reinterpret_cast<const uint32_t*>(O + ((int64_t)b * Mq + r) * ldo);
  const uint32_t* d = reinterpret_cast<const uint32_t*>(dO + ((int64_t)b * Mq + r) * ldo);
  for (int h = 0; h < H; ++h) {
    const uint32_t a = o[h * 32 + lane], g = d[h * 32 + lane];
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&a));
    const float2 fg = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g));
    const float s = warp_sum(fa.x * fg.x + fa.y * fg.y);
    if (lane == 0) {
      sc.delta[((int64_t)b * H + h) * S + r] = -s * (rs * kLn2);
      sc.nlse[((int64_t)b * H + h) * S + r] = -lse2[((int64_t)b * H + h) * S + r];
    }
  }
}
__global__ void attn_blk_kernel(int B, int S, Scratch sc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (S / 64)) return;
  int lo = INT_MAX, hi = INT_MIN, lo_max = INT_MIN, hi_min = INT_MAX;
  for (int r = 0; r < 64; ++r) {
    const int l = sc.row_lo[(int64_t)i * 64 + r], h = sc.row_hi[(int64_t)i * 64 + r];
    if (h > l) { lo = min(lo, l); hi = max(hi, h); }
    const bool normal = (h > l) && sc.row_scale[(int64_t)i * 64 + r] != 0.f;
    lo_max = normal ? max(lo_max, l) : INT_MAX;
    hi_min = normal ? min(hi_min, h) : INT_MIN;
  }
  sc.blk_lo[i] = lo;
  sc.blk_hi[i] = hi;
  sc.blk_lo_max[i] = lo_max;
  sc.blk_hi_min[i] = hi_min;
}

// ------------------------------------------------------------------------------------------------ dQ
struct DqParams {
  int B, H, Mq, Nk, S;
  Scratch sc;
  uint16_t* dQ;
  int64_t lddq;
};
constexpr int kDqStages = 3;
struct DqSmem {
  static constexpr int kQ = 0, kDO = kQ + kT * 128, kK = kDO + kT * 128, kV = kK + kDqStages * kBlk * 128,
                       kDS = kV + kDqStages * kBlk * 128, kBar = kDS + kT * 128, kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kThreads, 2)
attn_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
               const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const DqParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *sQ = smem + DqSmem::kQ, *sDO = smem + DqSmem::kDO, *sK = smem + DqSmem::kK, *sV = smem + DqSmem::kV,
          *sDS = smem + DqSmem::kDS;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int q0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kDqStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 128);
    mbar_init(ds_full, 128);
    mbar_init(ds_empty, 1);
    mbar_init(dq_full, 1);
    fence_mbar_init();
    s_range[0] = INT_MAX;
    s_range[1] = INT_MIN;
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  __syncthreads();
  const int row = q0 + tid;
  int lo = INT_MAX, hi = INT_MIN;
  float rscale = 0.f;
  if (warp < 4 && row < p.Mq) {
    lo = p.sc.row_lo[(int64_t)b * p.S + row];
    hi = p.sc.row_hi[(int64_t)b * p.S + row];
    rscale = p.sc.row_scale[(int64_t)b * p.S + row];
    if (hi > lo) { atomicMin(&s_range[0], lo); atomicMax(&s_range[1], hi); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lo_cta = s_range[0], hi_cta = s_range[1];
  const int nblk = hi_cta > lo_cta ? (hi_cta - lo_cta + kBlk - 1) / kBlk : 0;

  if (warp == 4) {
    if (lane == 0 && nblk > 0) {
      mbar_expect_tx(q_full, 2 * kT * 128);
      tma_load_2d(sQ, &tmQ, q_full, h * kD, b * p.Mq + q0);
      tma_load_2d(sDO, &tmDO, q_full, h * kD, b * p.Mq + q0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kDqStages;
        mbar_wait(&kv_empty[st], ((j / kDqStages) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * kBlk * 128);
        const int krow = b * p.Nk + lo_cta + j * kBlk;
        tma_load_2d(sK + st * kBlk * 128, &tmK, &kv_full[st], h * kD, krow);
        tma_load_2d(sV + st * kBlk * 128, &tmV, &kv_full[st], h * kD, krow);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBlk, 0, 0);
      constexpr uint32_t idesc_kmn = umma_idesc_bf16(128, kD, 0, 1);
      const uint32_t tS = tmem_base, tDP = tmem_base + 64, tDQ = tmem_base + 128;
      const uint32_t aQ = smem_u32(sQ), aDO = smem_u32(sDO), aDS = smem_u32(sDS);
      auto issue_sdp = [&](int st) {
        const uint32_t aK = smem_u32(sK + st * kBlk * 128), aV = smem_u32(sV + st * kBlk * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, umma_desc_kmajor_sw128(aQ + k * 32), umma_desc_kmajor_sw128(aK + k * 32), idesc_kk, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDP, umma_desc_kmajor_sw128(aDO + k * 32), umma_desc_kmajor_sw128(aV + k * 32), idesc_kk, k ? 1u : 0u);
        umma_commit(sdp_full);
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
        const uint32_t aK = smem_u32(sK + st * kBlk * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDQ, umma_desc_kmajor_sw128(aDS + k * 32), umma_desc_mnmajor_sw128(aK + k * 2048, 8192), idesc_kmn,
                       (j | k) ? 1u : 0u);
        umma_commit(ds_empty);
        umma_commit(&kv_empty[st]);
      }
      umma_commit(dq_full);
    }
  } else {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float nlse = -INFINITY, ndelta = 0.f;   // -lse2 and -delta * rs_nat
    if (row < p.Mq) {
      nlse = p.sc.nlse[((int64_t)b * p.H + h) * p.S + row];
      ndelta = p.sc.delta[((int64_t)b * p.H + h) * p.S + row];
    }
    const float rs_nat = rscale * kLn2;
    for (int j = 0; j < nblk; ++j) {
      const int kv0 = lo_cta + j * kBlk;
      mbar_wait(sdp_full, j & 1);
      tc_fence_after();
      uint32_t s0[32], s1[32], d0[32], d1[32];
      tmem_ld32(tmem_base + lane_addr, s0);
      tmem_ld32(tmem_base + lane_addr + 32, s1);
      tmem_ld32(tmem_base + lane_addr + 64, d0);
      tmem_ld32(tmem_base + lane_addr + 96, d1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(sdp_free);
      float sc = rscale;
      if (!(rscale != 0.f && kv0 >= lo && kv0 + kBlk <= hi)) {
        sc = rscale != 0.f ? rscale : 1.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const bool ok0 = (kv0 + c >= lo) && (kv0 + c < hi), ok1 = (kv0 + 32 + c >= lo) && (kv0 + 32 + c < hi);
          s0[c] = ok0 ? (rscale != 0.f ? s0[c] : 0u) : 0xff800000u;
          s1[c] = ok1 ? (rscale != 0.f ? s1[c] : 0u) : 0xff800000u;
        }
      }
      if (j > 0) mbar_wait(ds_empty, (j - 1) & 1);
#pragma unroll
      for (int c8 = 0; c8 < kBlk / 8; ++c8) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c8 * 8 + u;
          const float s = __uint_as_float(c < 32 ? s0[c] : s1[c - 32]);
          const float dp = __uint_as_float(c < 32 ? d0[c] : d1[c - 32]);
          e[u] = ex2(fmaf(s, sc, nlse)) * fmaf(dp, rs_nat, ndelta);
        }
        uint4 pk;
        pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
        *reinterpret_cast<uint4*>(sDS + swz_off(tid, c8)) = pk;
      }
      fence_async_smem();
      mbar_arrive(ds_full);
    }
    uint32_t v0[32], v1[32];
    if (nblk > 0) {
      mbar_wait(dq_full, 0);
      tc_fence_after();
      tmem_ld32(tmem_base + lane_addr + 128, v0);
      tmem_ld32(tmem_base + lane_addr + 160, v1);
      tmem_ld_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) { v0[c] = 0u; v1[c] = 0u; }
    }
    if (row < p.Mq) {
      uint16_t* out = p.dQ + ((int64_t)b * p.Mq + row) * p.lddq + h * kD;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint32_t* src = c8 < 4 ? &v0[c8 * 8] : &v1[(c8 - 4) * 8];
        uint4 pk;
        pk.x = pack_bf16(__uint_as_float(src[0]), __uint_as_float(src[1]));
        pk.y = pack_bf16(__uint_as_float(src[2]), __uint_as_float(src[3]));
        pk.z = pack_bf16(__uint_as_float(src[4]), __uint_as_float(src[5]));
        pk.w = pack_bf16(__uint_as_float(src[6]), __uint_as_float(src[7]));
        reinterpret_cast<uint4*>(out)[c8] = pk;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
struct DkvParams {
  int B, H, Mq, Nk, S;
  float scale_log2;
  Scratch sc;
  uint16_t* dK;
  uint16_t* dV;
  int64_t lddk, lddv;
};
constexpr int kDkvStages = 2;
constexpr int kMetaBytes = 5 * kBlk * 4;  // -lse2, -delta*rs, lo, hi, scale for 64 query rows
constexpr int kMaxQBlocks = 1024;
struct DkvSmem {
  static constexpr int kK = 0, kV = kK + kT * 128, kQ = kV + kT * 128, kDO = kQ + kDkvStages * kBlk * 128,
                       kPT = kDO + kDkvStages * kBlk * 128, kDST = kPT + kT * 128, kMeta = kDST + kT * 128,
                       kList = kMeta + kDkvStages * kMetaBytes, kBar = kList + kMaxQBlocks * 2, kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kThreads, 2)
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int kv0 = blockIdx.x * kT, h = blockIdx.y, b = blockIdx.z;
  const int nqb = p.S / kBlk;
  if (tid == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < kDkvStages; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(sdp_full, 1);
    mbar_init(sdp_free, 128);
    mbar_init(pds_full, 128);
    mbar_init(pds_empty, 1);
    mbar_init(dkv_full, 1);
    fence_mbar_init();
    int n = 0;  // query blocks whose key ranges intersect this key tile
    for (int i = 0; i < nqb; ++i) {
      const int64_t o = (int64_t)b * nqb + i;
      const int bl = p.sc.blk_lo[o], bh = p.sc.blk_hi[o];
      if (bh > kv0 && bl < kv0 + kT && n < kMaxQBlocks) {
        const bool inside = p.sc.blk_lo_max[o] <= kv0 && p.sc.blk_hi_min[o] >= kv0 + kT;
        s_list[n++] = (uint16_t)(i | (inside ? 0x8000 : 0));
      }
    }
    *s_n = n;
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n = *s_n;

  if (warp == 4) {
    if (lane == 0 && n > 0) {
      mbar_expect_tx(kv_full, 2 * kT * 128);
      tma_load_2d(sK, &tmK, kv_full, h * kD, b * p.Nk + kv0);
      tma_load_2d(sV, &tmV, kv_full, h * kD, b * p.Nk + kv0);
      for (int idx = 0; idx < n; ++idx) {
        const int st = idx % kDkvStages;
        const int r0 = (int)(s_list[idx] & 0x7fff) * kBlk;
        mbar_wait(&q_empty[st], ((idx / kDkvStages) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * kBlk * 128 + kMetaBytes);
        tma_load_2d(sQ + st * kBlk * 128, &tmQ, &q_full[st], h * kD, b * p.Mq + r0);
        tma_load_2d(sDO + st * kBlk * 128, &tmDO, &q_full[st], h * kD, b * p.Mq + r0);
        uint8_t* meta = sMeta + st * kMetaBytes;
        const int64_t hoff = ((int64_t)b * p.H + h) * p.S + r0, roff = (int64_t)b * p.S + r0;
        bulk_load(meta + 0 * 256, p.sc.nlse + hoff, 256, &q_full[st]);
        bulk_load(meta + 1 * 256, p.sc.delta + hoff, 256, &q_full[st]);
        bulk_load(meta + 2 * 256, p.sc.row_lo + roff, 256, &q_full[st]);
        bulk_load(meta + 3 * 256, p.sc.row_hi + roff, 256, &q_full[st]);
        bulk_load(meta + 4 * 256, p.sc.row_scale + roff, 256, &q_full[st]);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && n > 0) {
      constexpr uint32_t idesc_kk = umma_idesc_bf16(128, kBlk, 0, 0);
      constexpr uint32_t idesc_kmn = umma_idesc_bf16(128, kD, 0, 1);
      const uint32_t tST = tmem_base, tDPT = tmem_base + 64, tDV = tmem_base + 128, tDK = tmem_base + 192;
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aPT = smem_u32(sPT), aDST = smem_u32(sDST);
      auto issue_sdp = [&](int st) {
        const uint32_t aQ = smem_u32(sQ + st * kBlk * 128), aDO = smem_u32(sDO + st * kBlk * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tST, umma_desc_kmajor_sw128(aK + k * 32), umma_desc_kmajor_sw128(aQ + k * 32), idesc_kk, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDPT, umma_desc_kmajor_sw128(aV + k * 32), umma_desc_kmajor_sw128(aDO + k * 32), idesc_kk, k ? 1u : 0u);
        umma_commit(sdp_full);
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
        const uint32_t aQ = smem_u32(sQ + st * kBlk * 128), aDO = smem_u32(sDO + st * kBlk * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDV, umma_desc_kmajor_sw128(aPT + k * 32), umma_desc_mnmajor_sw128(aDO + k * 2048, 8192), idesc_kmn,
                       (idx | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tDK, umma_desc_kmajor_sw128(aDST + k * 32), umma_desc_mnmajor_sw128(aQ + k * 2048, 8192), idesc_kmn,
                       (idx | k) ? 1u : 0u);
        umma_commit(pds_empty);
        umma_commit(&q_empty[st]);
      }
      umma_commit(dkv_full);
    }
  } else {
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const int kidx = kv0 + tid;
    const float SC = p.scale_log2, RN = p.scale_log2 * kLn2;
    for (int idx = 0; idx < n; ++idx) {
      const int st = idx % kDkvStages;
      const bool inside = (s_list[idx] & 0x8000) != 0;
      mbar_wait(&q_full[st], (idx / kDkvStages) & 1);  // row metadata visible
      mbar_wait(sdp_full, idx & 1);
      tc_fence_after();
      const float* m_nl = reinterpret_cast<const float*>(sMeta + st * kMetaBytes);
      const float* m_nd = m_nl + 64;
      const int* m_lo = reinterpret_cast<const int*>(m_nl + 128);
      const int* m_hi = m_lo + 64;
      const float* m_rs = m_nl + 256;
      uint32_t s0[32], s1[32], d0[32], d1[32];
      tmem_ld32(tmem_base + lane_addr, s0);
      tmem_ld32(tmem_base + lane_addr + 32, s1);
      tmem_ld32(tmem_base + lane_addr + 64, d0);
      tmem_ld32(tmem_base + lane_addr + 96, d1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(sdp_free);
      if (idx > 0) mbar_wait(pds_empty, (idx - 1) & 1);
      if (inside) {
        // every (key, query) pair of this block is unmasked and every row uses the plain scale
#pragma unroll
        for (int c8 = 0; c8 < kBlk / 8; ++c8) {
          const float4 nl0 = *reinterpret_cast<const float4*>(m_nl + c8 * 8), nl1 = *reinterpret_cast<const float4*>(m_nl + c8 * 8 + 4);
          const float4 nd0 = *reinterpret_cast<const float4*>(m_nd + c8 * 8), nd1 = *reinterpret_cast<const float4*>(m_nd + c8 * 8 + 4);
          const float nl[8] = {nl0.x, nl0.y, nl0.z, nl0.w, nl1.x, nl1.y, nl1.z, nl1.w};
          const float nd[8] = {nd0.x, nd0.y, nd0.z, nd0.w, nd1.x, nd1.y, nd1.z, nd1.w};
          float pt[8], e[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c8 * 8 + u;
            const float s = __uint_as_float(c < 32 ? s0[c] : s1[c - 32]);
            const float dp = __uint_as_float(c < 32 ? d0[c] : d1[c - 32]);
            pt[u] = ex2(fmaf(s, SC, nl[u]));
            e[u] = pt[u] * fmaf(dp, RN, nd[u]);
          }
          uint4 pk;
          pk.x = pack_bf16(pt[0], pt[1]); pk.y = pack_bf16(pt[2], pt[3]); pk.z = pack_bf16(pt[4], pt[5]); pk.w = pack_bf16(pt[6], pt[7]);
          *reinterpret_cast<uint4*>(sPT + swz_off(tid, c8)) = pk;
          pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(sDST + swz_off(tid, c8)) = pk;
        }
      } else {
#pragma unroll
        for (int c8 = 0; c8 < kBlk / 8; ++c8) {
          float pt[8], e[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int c = c8 * 8 + u;
            const float s = __uint_as_float(c < 32 ? s0[c] : s1[c - 32]);
            const float dp = __uint_as_float(c < 32 ? d0[c] : d1[c - 32]);
            const float rs = m_rs[c];
            const bool ok = kidx >= m_lo[c] && kidx < m_hi[c];
            pt[u] = ok ? ex2(fmaf(rs != 0.f ? s : 0.f, rs, m_nl[c])) : 0.f;
            e[u] = pt[u] * fmaf(dp, rs * kLn2, m_nd[c]);
          }
          uint4 pk;
          pk.x = pack_bf16(pt[0], pt[1]); pk.y = pack_bf16(pt[2], pt[3]); pk.z = pack_bf16(pt[4], pt[5]); pk.w = pack_bf16(pt[6], pt[7]);
          *reinterpret_cast<uint4*>(sPT + swz_off(tid, c8)) = pk;
          pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(sDST + swz_off(tid, c8)) = pk;
        }
      }
      fence_async_smem();
      mbar_arrive(pds_full);
    }
    if (n > 0) {
      mbar_wait(dkv_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {  // 0: dV (cols 128..191), 1: dK (cols 192..255)
      uint32_t v0[32], v1[32];
      if (n > 0) {
        tmem_ld32(tmem_base + lane_addr + 128 + which * 64, v0);
        tmem_ld32(tmem_base + lane_addr + 160 + which * 64, v1);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) { v0[c] = 0u; v1[c] = 0u; }
      }
      if (kidx < p.Nk) {
        uint16_t* out = which == 0 ? p.dV + ((int64_t)b * p.Nk + kidx) * p.lddv + h * kD
                                   : p.dK + ((int64_t)b * p.Nk + kidx) * p.lddk + h * kD;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          const uint32_t* src = c8 < 4 ? &v0[c8 * 8] : &v1[(c8 - 4) * 8];
          uint4 pk;
          pk.x = pack_bf16(__uint_as_float(src[0]), __uint_as_float(src[1]));
          pk.y = pack_bf16(__uint_as_float(src[2]), __uint_as_float(src[3]));
          pk.z = pack_bf16(__uint_as_float(src[4]), __uint_as_float(src[5]));
          pk.w = pack_bf16(__uint_as_float(src[6]), __uint_as_float(src[7]));
          reinterpret_cast<uint4*>(out)[c8] = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace egom2p

extern "C" int64_t egom2p_attn_bwd_scratch_bytes(int32_t B, int32_t H, int32_t Mq) {
  const int64_t S = egom2p::pad64(Mq);
  return 2 * (int64_t)B * H * S * 4 + 3 * (int64_t)B * S * 4 + 4 * ((int64_t)B * (S / 64) * 4 + 512) + 1024;
}

extern "C" int egom2p_attn_bwd(const uint16_t* Q, const uint16_t* K, const uint16_t* V, const uint16_t* O, const uint16_t* dO,
                               const float* lse, int32_t B, int32_t H, int32_t Mq, int32_t Nk, int64_t ldq, int64_t ldk,
                               int64_t ldv, int64_t ldo, const int32_t* key_lo, const int32_t* key_hi, float scale,
                               void* scratch, uint16_t* dQ, uint16_t* dK, uint16_t* dV, int64_t lddq, int64_t lddk,
                               int64_t lddv, void* stream_) {
  using namespace egom2p;
  cudaStream_t stream = (cudaStream_t)stream_;
  EGO_REQUIRE(Q && O && dO && lse && scratch && dQ && B > 0 && H > 0 && Mq > 0 && Nk >= 0, "attn_bwd: bad argument");
  EGO_REQUIRE((key_lo == nullptr) == (key_hi == nullptr), "attn_bwd: key_lo / key_hi must both be given or both NULL");
  EGO_REQUIRE(((uintptr_t)scratch & 255) == 0, "attn_bwd: scratch must be 256-byte aligned");
  EGO_REQUIRE(lddq % 8 == 0 && ((uintptr_t)dQ & 15) == 0, "attn_bwd: dQ alignment");
  const int S = pad64(Mq);
  Scratch sc = carve(scratch, B, H, Mq);
  attn_prep_kernel<<<(unsigned)(((int64_t)B * S + 7) / 8), 256, 0, stream>>>(O, dO, ldo, B, H, Mq, Nk, S, key_lo, key_hi, lse,
                                                                          scale * kLog2e, sc);
  int rc = check_launch("attn_bwd prep");
  if (rc) return rc;
  attn_blk_kernel<<<(B * (S / 64) + 127) / 128, 128, 0, stream>>>(B, S, sc);
  if ((rc = check_launch("attn_bwd blk"))) return rc;

  CUtensorMap tmQ, tmDO, tmK, tmV;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem::kTotal);
    cudaError_t e2 = cudaFuncSetAttribute(attn_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem::kTotal);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_bwd: cudaFuncSetAttribute failed"); return EGOM2P_ERR_CUDA; }
    attr_set = true;
  }
  // ---- dQ
  if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kT, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kT, kD))) return rc;
  if (Nk > 0) {
    EGO_REQUIRE(K && V && dK && dV, "attn_bwd: K / V / dK / dV missing");
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kBlk, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kBlk, kD))) return rc;
  } else {
    tmK = tmQ; tmV = tmQ;
  }
  DqParams pq{B, H, Mq, Nk, S, sc, dQ, lddq};
  attn_dq_kernel<<<dim3((Mq + kT - 1) / kT, H, B), kThreads, DqSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, pq);
  if ((rc = check_launch("attn_bwd dq"))) return rc;
  if (Nk == 0) return EGOM2P_OK;
  // ---- dK / dV
  EGO_REQUIRE(S / kBlk <= kMaxQBlocks, "attn_bwd: Mq too large (max %d)", kMaxQBlocks * kBlk);
  EGO_REQUIRE(lddk % 8 == 0 && lddv % 8 == 0 && ((uintptr_t)dK & 15) == 0 && ((uintptr_t)dV & 15) == 0, "attn_bwd: dK/dV alignment");
  if ((rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kBlk, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmDO, dO, (uint64_t)B * Mq, (uint64_t)H * kD, ldo, kBlk, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kT, kD))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kT, kD))) return rc;
  DkvParams pk{B, H, Mq, Nk, S, scale * kLog2e, sc, dK, dV, lddk, lddv};
  attn_dkv_kernel<<<dim3((Nk + kT - 1) / kT, H, B), kThreads, DkvSmem::kTotal, stream>>>(tmQ, tmDO, tmK, tmV, pk);
  return check_launch("attn_bwd dkv");
}
