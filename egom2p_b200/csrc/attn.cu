// Range-masked flash attention for head_dim 64 on tcgen05/TMEM, operands staged by TMA (north-star kernel 2).
//
// Every mask the path produces is a contiguous key range per query row: encoder self / decoder
// cross = [0, n_enc[b]); decoder self = the row's own modality segment. Rows with an empty range reproduce the
// reference's masked_fill(-finfo.max) semantics (uniform attention over all Nk keys, A4/A5).
//
// One CTA = 128 query rows of one (batch, head); 2 CTAs are co-resident per SM so one CTA's softmax overlaps the
// other's MMAs. Warps 0-3: softmax / gradient math (thread == query row == TMEM lane); warp 4: TMA producer;
// warp 5: tcgen05.mma issuer + TMEM owner. Key blocks are 64 wide so a score row lives in 64 registers.
//
//   fwd : S = Q K^T (TMEM) -> online softmax -> P (bf16, swizzled smem) -> O_blk = P V (TMEM) -> O += in registers
//   dQ  : S, dP = dO V^T (TMEM) -> dS (smem) -> dQ += dS K   (accumulated in TMEM over the key loop)
//   dKV : S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (smem) -> dV += P^T dO, dK += dS^T Q (TMEM accumulators)
#include <climits>

#include "common.cuh"

namespace egom2p {

constexpr int kAttnThreads = 192;
constexpr int kD = 64;       // head dim
constexpr int kQT = 128;     // query rows per CTA (fwd, dQ) / key rows per CTA (dKV)
constexpr int kKB = 64;      // inner block (keys in fwd/dQ, query rows in dKV)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {  // single MUFU.EX2 (exp2f adds range fix-ups the softmax does not need)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t swz_off(int row, int chunk) {  // byte offset of 16-byte chunk in a [rows][128 B] SW128 tile
  return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}

struct AttnFwdParams {
  int B, H, Mq, Nk, Mq_pad;
  const int32_t* key_lo;
  const int32_t* key_hi;
  float scale_log2;
  uint16_t* O;
  int64_t ldo;
  float* lse2;  // (B, H, Mq_pad), log2 domain: m + log2(l)
};

constexpr int kFwdStages = 4;
struct FwdSmem {
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kQT * 128;
  static constexpr int kV = kK + kFwdStages * kKB * 128;
  static constexpr int kP = kV + kFwdStages * kKB * 128;
  static constexpr int kBar = kP + kQT * 128;
  static constexpr int kTotal = kBar + 256 + 1024;
};

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + FwdSmem::kQ;
  uint8_t* sK = smem + FwdSmem::kK;
  uint8_t* sV = smem + FwdSmem::kV;
  uint8_t* sP = smem + FwdSmem::kP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBar);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kFwdStages;
  uint64_t* s_full = kv_empty + kFwdStages;
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = p_full + 1;
  uint64_t* s_free = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 1);
  int* s_range = reinterpret_cast<int*>(tmem_slot + 1);  // [0] = lo, [1] = hi

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int q0 = blockIdx.x * kQT, h = blockIdx.y, b = blockIdx.z;

  if (tid == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < kFwdStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(s_free, 128);
    fence_mbar_init();
    s_range[0] = INT_MAX;
    s_range[1] = INT_MIN;
  }
  if (warp == 5) tmem_alloc<128>(tmem_slot);
  __syncthreads();

  // per-row key range
  int lo = INT_MAX, hi = INT_MIN;
  float rscale = p.scale_log2;
  const int row = q0 + tid;
  if (warp < 4 && row < p.Mq) {
    lo = p.key_lo ? p.key_lo[(int64_t)b * p.Mq + row] : 0;
    hi = p.key_hi ? p.key_hi[(int64_t)b * p.Mq + row] : p.Nk;
    hi = min(hi, p.Nk);
    lo = max(lo, 0);
    if (hi <= lo) { lo = 0; hi = p.Nk; rscale = 0.f; }  // every key masked -> uniform over all keys
    if (hi > lo) { atomicMin(&s_range[0], lo); atomicMax(&s_range[1], hi); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int lo_cta = s_range[0], hi_cta = s_range[1];
  const int nblk = hi_cta > lo_cta ? (hi_cta - lo_cta + kKB - 1) / kKB : 0;

  if (warp == 4) {
    if (lane == 0 && nblk > 0) {
      mbar_expect_tx(q_full, kQT * 128);
      tma_load_2d(sQ, &tmQ, q_full, h * kD, b * p.Mq + q0);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kFwdStages;
        mbar_wait(&kv_empty[st], ((j / kFwdStages) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * kKB * 128);
        const int krow = b * p.Nk + lo_cta + j * kKB;
        tma_load_2d(sK + st * kKB * 128, &tmK, &kv_full[st], h * kD, krow);
        tma_load_2d(sV + st * kKB * 128, &tmV, &kv_full[st], h * kD, krow);
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && nblk > 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kKB, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, kD, 0, 1);
      const uint32_t tS = tmem_base, tO = tmem_base + 64;
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16_ss(tS, umma_desc_kmajor_sw128(aQ + k * 32), umma_desc_kmajor_sw128(smem_u32(sK) + k * 32), idesc_s, k ? 1u : 0u);
      umma_commit(s_full);
      for (int j = 0; j < nblk; ++j) {
        const int st = j % kFwdStages;
        if (j + 1 < nblk) {  // S(j+1) as soon as S(j) sits in registers: overlaps the softmax of block j
          const int st1 = (j + 1) % kFwdStages;
          mbar_wait(s_free, j & 1);
          mbar_wait(&kv_full[st1], ((j + 1) / kFwdStages) & 1);
          tc_fence_after();
          const uint32_t aK = smem_u32(sK + st1 * kKB * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_ss(tS, umma_desc_kmajor_sw128(aQ + k * 32), umma_desc_kmajor_sw128(aK + k * 32), idesc_s, k ? 1u : 0u);
          umma_commit(s_full);
        }
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + st * kKB * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tO, umma_desc_kmajor_sw128(aP + k * 32), umma_desc_mnmajor_sw128(aV + k * 2048, 8192), idesc_pv, k ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warps: thread == query row
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o[kD];
#pragma unroll
    for (int c = 0; c < kD; ++c) o[c] = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int kv0 = lo_cta + j * kKB;
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_base + lane_addr, v0);
      tmem_ld32(tmem_base + lane_addr + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      float sc = rscale;
      if (!(rscale != 0.f && kv0 >= lo && kv0 + kKB <= hi)) {  // block touches the range boundary (or uniform row)
        sc = rscale != 0.f ? rscale : 1.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const bool ok0 = (kv0 + c >= lo) && (kv0 + c < hi), ok1 = (kv0 + 32 + c >= lo) && (kv0 + 32 + c < hi);
          v0[c] = ok0 ? (rscale != 0.f ? v0[c] : 0u) : 0xff800000u;
          v1[c] = ok1 ? (rscale != 0.f ? v1[c] : 0u) : 0xff800000u;
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY, bm2 = -INFINITY, bm3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        bm0 = fmaxf(bm0, fmaxf(__uint_as_float(v0[c]), __uint_as_float(v1[c])));
        bm1 = fmaxf(bm1, fmaxf(__uint_as_float(v0[c + 1]), __uint_as_float(v1[c + 1])));
        bm2 = fmaxf(bm2, fmaxf(__uint_as_float(v0[c + 2]), __uint_as_float(v1[c + 2])));
        bm3 = fmaxf(bm3, fmaxf(__uint_as_float(v0[c + 3]), __uint_as_float(v1[c + 3])));
      }
      const float bm = fmaxf(fmaxf(bm0, bm1), fmaxf(bm2, bm3)) * sc;  // sc > 0, so max commutes with the scaling
      const float m_new = fmaxf(m, bm);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2(m - m_use);
      const float nm = -m_use;
      float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < kKB / 8; ++c8) {
        float e[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = c8 * 8 + u;
          e[u] = ex2(fmaf(__uint_as_float(c < 32 ? v0[c] : v1[c - 32]), sc, nm));
        }
        sum0 += e[0] + e[4]; sum1 += e[1] + e[5]; sum2 += e[2] + e[6]; sum3 += e[3] + e[7];
        uint4 pk;
        pk.x = pack_bf16(e[0], e[1]); pk.y = pack_bf16(e[2], e[3]); pk.z = pack_bf16(e[4], e[5]); pk.w = pack_bf16(e[6], e[7]);
        *reinterpret_cast<uint4*>(sP + swz_off(tid, c8)) = pk;
      }
      const float sum = (sum0 + sum1) + (sum2 + sum3);
      l = l * alpha + sum;
      m = m_new;
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      tmem_ld32(tmem_base + lane_addr + 64, v0);
      tmem_ld32(tmem_base + lane_addr + 96, v1);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        o[c] = o[c] * alpha + __uint_as_float(v0[c]);
        o[c + 32] = o[c + 32] * alpha + __uint_as_float(v1[c]);
      }
    }
    if (row < p.Mq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      uint16_t* orow = p.O + ((int64_t)b * p.Mq + row) * p.ldo + h * kD;
#pragma unroll
      for (int c8 = 0; c8 < kD / 8; ++c8) {
        uint4 pk;
        pk.x = pack_bf16(o[c8 * 8 + 0] * inv, o[c8 * 8 + 1] * inv);
        pk.y = pack_bf16(o[c8 * 8 + 2] * inv, o[c8 * 8 + 3] * inv);
        pk.z = pack_bf16(o[c8 * 8 + 4] * inv, o[c8 * 8 + 5] * inv);
        pk.w = pack_bf16(o[c8 * 8 + 6] * inv, o[c8 * 8 + 7] * inv);
        reinterpret_cast<uint4*>(orow)[c8] = pk;
      }
      if (p.lse2) p.lse2[((int64_t)b * p.H + h) * p.Mq_pad + row] = l > 0.f ? m + log2f(l) : INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

}  // namespace egom2p

extern "C" int egom2p_attn_lse_stride(int32_t Mq) { return (Mq + 63) / 64 * 64; }

extern "C" int egom2p_attn_fwd(const uint16_t* Q, const uint16_t* K, const uint16_t* V, int32_t B, int32_t H, int32_t Mq,
                               int32_t Nk, int64_t ldq, int64_t ldk, int64_t ldv, const int32_t* key_lo,
                               const int32_t* key_hi, float scale, uint16_t* O, int64_t ldo, float* lse, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(Q && O && B > 0 && H > 0 && Mq > 0 && Nk >= 0, "attn_fwd: bad argument");
  EGO_REQUIRE(ldo % 8 == 0 && ((uintptr_t)O & 15) == 0, "attn_fwd: O must be 16-byte aligned with ldo %% 8 == 0");
  EGO_REQUIRE((key_lo == nullptr) == (key_hi == nullptr), "attn_fwd: key_lo / key_hi must both be given or both NULL");
  AttnFwdParams p{};
  p.B = B; p.H = H; p.Mq = Mq; p.Nk = Nk; p.Mq_pad = egom2p_attn_lse_stride(Mq);
  p.key_lo = key_lo; p.key_hi = key_hi; p.scale_log2 = scale * kLog2e; p.O = O; p.ldo = ldo; p.lse2 = lse;
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_bf16_2d(&tmQ, Q, (uint64_t)B * Mq, (uint64_t)H * kD, ldq, kQT, kD);
  if (rc) return rc;
  if (Nk > 0) {
    EGO_REQUIRE(K && V, "attn_fwd: K / V missing");
    if ((rc = make_tmap_bf16_2d(&tmK, K, (uint64_t)B * Nk, (uint64_t)H * kD, ldk, kKB, kD))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmV, V, (uint64_t)B * Nk, (uint64_t)H * kD, ldv, kKB, kD))) return rc;
  } else {
    tmK = tmQ;
    tmV = tmQ;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal);
    if (e != cudaSuccess) { set_error("attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return EGOM2P_ERR_CUDA; }
    attr_set = true;
  }
  dim3 grid((Mq + kQT - 1) / kQT, H, B);
  attn_fwd_kernel<<<grid, kAttnThreads, FwdSmem::kTotal, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  return check_launch("attn_fwd");
}

// ---- backward entry points (implemented below in attn_bwd.cu once built; placeholders keep the ABI complete)
