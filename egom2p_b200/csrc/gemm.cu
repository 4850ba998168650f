// bf16 GEMM on the 5th-gen tensor cores: tcgen05.mma (cta_group::1, 128 x BN x 16) with fp32 accumulators in TMEM,
// operands staged by TMA (128B swizzle) through an mbarrier ring, persistent over output tiles with a double-buffered
// accumulator so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue
// (TMEM -> registers -> per-warp smem transpose -> coalesced global stores with fused bias / residual add, or the
// cross-entropy epilogues that never write logits to HBM).
//
// Operand majors (see include/egom2p_b200.h): K-major tiles are [rows][64 k] (one TMA box); MN-major tiles are
// [64 k][64 mn] boxes, one per 64 rows of the tile, consumed through MN-major UMMA descriptors -- this is what lets
// dgrad (B = W stored [N][K]) and wgrad (A = dY^T, B = X^T) run without any transposed copies.
#include "common.cuh"

namespace egom2p {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;
constexpr int kStagePitch = 36;  // floats; 16-byte aligned rows, conflict-free float4 phases

enum { EPI_STORE = 0, EPI_CE_PARTIAL = 1, EPI_CE_DLOGITS = 2 };

struct GemmParams {
  int M, N, K;
  // EPI_STORE
  const float* bias;
  const float* addend;
  int64_t ld_add;
  uint16_t* c_bf16;
  float* c_f32;
  int64_t ldc;
  // CE epilogues
  const int64_t* target;
  const float* lse;
  const float* gscale;
  int v0;
  float* part_max;
  float* part_sum;
  float* tgt_logit;
};

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingBytes = 4 * 32 * kStagePitch * 4;
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kTotal = kStages * kStageBytes + kStagingBytes + kBarBytes + 1024;  // + alignment slack
};

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::kABytes;
  float* staging = reinterpret_cast<float*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes + S::kStagingBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int total_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const int k0 = kb * BK;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          uint8_t* a = sA + stage * S::kABytes;
          uint8_t* b = sB + stage * S::kBBytes;
          if (!A_MN) {
            tma_load_2d(a, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int h = 0; h < BM / 64; ++h) tma_load_2d(a + h * 8192, &tmA, &full_bar[stage], m0 + h * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d(b, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h) tma_load_2d(b + h * 8192, &tmB, &full_bar[stage], n0 + h * 64, k0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * S::kABytes);
          const uint32_t b_addr = smem_u32(sB + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc_mnmajor_sw128(a_addr + k * 2048, 8192) : umma_desc_kmajor_sw128(a_addr + k * 32);
            const uint64_t db = B_MN ? umma_desc_mnmajor_sw128(b_addr + k * 2048, 8192) : umma_desc_kmajor_sw128(b_addr + k * 32);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* st = staging + (warp - 2) * 32 * kStagePitch;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int n_blk = tile % n_tiles;
      const int m0 = (tile / n_tiles) * BM, n0 = n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int my_row = m0 + q * 32 + lane;
      float row_lse = 0.f, gs = 0.f;
      int row_tgt = -1;
      float run_max = -INFINITY, run_sum = 0.f, tl = 0.f;
      bool has_tl = false;
      if (EPI != EPI_STORE && my_row < p.M) {
        row_tgt = (int)p.target[my_row] - p.v0;
        if (EPI == EPI_CE_DLOGITS) {
          row_lse = p.lse[my_row];
          gs = *p.gscale;
        }
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int cbase = n0 + c * 32;
        if (EPI == EPI_CE_PARTIAL) {
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = (cbase + j < p.N) ? __uint_as_float(v[j]) : -INFINITY;
            cm = fmaxf(cm, f);
            if (cbase + j == row_tgt) { tl = f; has_tl = true; }
          }
          if (cm > -INFINITY) {
            const float nm = fmaxf(run_max, cm);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (cbase + j < p.N) s += __expf(__uint_as_float(v[j]) - nm);
            run_sum = run_sum * __expf(run_max - nm) + s;
            run_max = nm;
          }
          continue;
        }
        if (EPI == EPI_CE_DLOGITS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float pr = __expf(__uint_as_float(v[j]) - row_lse);
            if (cbase + j == row_tgt) pr -= 1.f;
            v[j] = __float_as_uint(pr * gs);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<uint4*>(st + lane * kStagePitch + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        const int colv = 4 * (lane & 7);
        const int gc = cbase + colv;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = (lane >> 3) + 4 * i;
          const int gr = m0 + q * 32 + r;
          if (gr < p.M && gc < p.N) {
            float4 o = *reinterpret_cast<const float4*>(st + r * kStagePitch + colv);
            if (p.bias) {
              const float4 bv = *reinterpret_cast<const float4*>(p.bias + gc);
              o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
            }
            if (p.addend) {
              const float4 av = *reinterpret_cast<const float4*>(p.addend + (int64_t)gr * p.ld_add + gc);
              o.x += av.x; o.y += av.y; o.z += av.z; o.w += av.w;
            }
            if (p.c_f32) *reinterpret_cast<float4*>(p.c_f32 + (int64_t)gr * p.ldc + gc) = o;
            if (p.c_bf16) {
              uint2 pk;
              pk.x = pack_bf16(o.x, o.y);
              pk.y = pack_bf16(o.z, o.w);
              *reinterpret_cast<uint2*>(p.c_bf16 + (int64_t)gr * p.ldc + gc) = pk;
            }
          }
        }
        __syncwarp();
      }
      if (EPI == EPI_CE_PARTIAL && my_row < p.M) {
        p.part_max[(int64_t)n_blk * p.M + my_row] = run_max;
        p.part_sum[(int64_t)n_blk * p.M + my_row] = run_sum;
        if (has_tl) p.tgt_logit[my_row] = tl;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm(const uint16_t* A, const uint16_t* B, int64_t lda, int64_t ldb, const GemmParams& p, cudaStream_t stream) {
  using S = GemmSmem<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  if (!A_MN) rc = make_tmap_bf16_2d(&tmA, A, p.M, p.K, lda, BM, BK);
  else       rc = make_tmap_bf16_2d(&tmA, A, p.K, p.M, lda, BK, 64);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap_bf16_2d(&tmB, B, p.N, p.K, ldb, BN, BK);
  else       rc = make_tmap_bf16_2d(&tmB, B, p.K, p.N, ldb, BK, 64);
  if (rc) return rc;
  auto kern = gemm_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
    if (e != cudaSuccess) {
      set_error("gemm: cudaFuncSetAttribute(%d B smem): %s", S::kTotal, cudaGetErrorString(e));
      return EGOM2P_ERR_CUDA;
    }
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  static int cached_sms = 0;
  if (!cached_sms) cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev);
  if (cached_sms > 0) sms = cached_sms;
  const int tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < sms ? tiles : sms;
  kern<<<grid, kGemmThreads, S::kTotal, stream>>>(tmA, tmB, p);
  return check_launch("gemm_bf16");
}

template <int EPI>
static int dispatch_gemm(const uint16_t* A, const uint16_t* B, int64_t lda, int64_t ldb, int a_mn, int b_mn,
                         const GemmParams& p, cudaStream_t stream) {
  // BN = 256 halves B re-reads per tile; BN = 128 gives better wave quantisation on small problems.
  const int sms = 148;
  const int64_t tiles256 = (int64_t)((p.M + BM - 1) / BM) * ((p.N + 255) / 256);
  const bool use256 = (EPI != EPI_STORE) || (p.N >= 256 && tiles256 >= 2 * sms);
#define EGO_GEMM_CASE(BN_, AM, BMJ) return launch_gemm<BN_, AM, BMJ, EPI>(A, B, lda, ldb, p, stream)
  if constexpr (EPI != EPI_STORE) {  // CE epilogues: Y (K-major) x W (K-major)
    EGO_GEMM_CASE(256, false, false);
  } else if (use256) {
    if (!a_mn && !b_mn) EGO_GEMM_CASE(256, false, false);
    if (!a_mn && b_mn) EGO_GEMM_CASE(256, false, true);
    if (a_mn && b_mn) EGO_GEMM_CASE(256, true, true);
    EGO_GEMM_CASE(256, true, false);
  } else {
    if (!a_mn && !b_mn) EGO_GEMM_CASE(128, false, false);
    if (!a_mn && b_mn) EGO_GEMM_CASE(128, false, true);
    if (a_mn && b_mn) EGO_GEMM_CASE(128, true, true);
    EGO_GEMM_CASE(128, true, false);
  }
#undef EGO_GEMM_CASE
}

}  // namespace egom2p

extern "C" int egom2p_gemm_bf16(const uint16_t* A, const uint16_t* B, int32_t M, int32_t N, int32_t K, int64_t lda,
                                int64_t ldb, int32_t a_mn, int32_t b_mn, const float* bias, const float* addend,
                                int64_t ld_add, uint16_t* c_bf16, float* c_f32, int64_t ldc, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(A && B && M > 0 && N > 0 && K > 0, "gemm_bf16: null operand or empty shape (M=%d N=%d K=%d)", M, N, K);
  EGO_REQUIRE(c_bf16 || c_f32, "gemm_bf16: no output");
  EGO_REQUIRE(N % 4 == 0 && ldc % 4 == 0 && (!addend || ld_add % 4 == 0), "gemm_bf16: N, ldc, ld_add must be multiples of 4");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.addend = addend; p.ld_add = ld_add; p.c_bf16 = c_bf16; p.c_f32 = c_f32; p.ldc = ldc;
  return dispatch_gemm<EPI_STORE>(A, B, lda, ldb, a_mn, b_mn, p, (cudaStream_t)stream);
}

extern "C" int egom2p_ce_partials(const uint16_t* Y, const uint16_t* W, const int64_t* target, int32_t R, int32_t V, int32_t K,
                                  int64_t ldy, int64_t ldw, float* part_max, float* part_sum, float* tgt_logit, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(Y && W && target && part_max && part_sum && tgt_logit && R > 0 && V > 0 && K > 0, "ce_partials: bad argument");
  GemmParams p{};
  p.M = R; p.N = V; p.K = K; p.target = target; p.v0 = 0; p.part_max = part_max; p.part_sum = part_sum; p.tgt_logit = tgt_logit;
  return dispatch_gemm<EPI_CE_PARTIAL>(Y, W, ldy, ldw, 0, 0, p, (cudaStream_t)stream);
}

extern "C" int egom2p_ce_dlogits(const uint16_t* Y, const uint16_t* W, const int64_t* target, const float* lse,
                                 const float* gscale, int32_t R, int32_t v0, int32_t Vc, int32_t K, int64_t ldy, int64_t ldw,
                                 uint16_t* dlogits, int64_t ldd, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(Y && W && target && lse && gscale && dlogits && R > 0 && Vc > 0 && K > 0 && v0 >= 0, "ce_dlogits: bad argument");
  EGO_REQUIRE(Vc % 4 == 0 && ldd % 4 == 0, "ce_dlogits: Vc and ldd must be multiples of 4");
  GemmParams p{};
  p.M = R; p.N = Vc; p.K = K; p.target = target; p.lse = lse; p.gscale = gscale; p.v0 = v0; p.c_bf16 = dlogits; p.ldc = ldd;
  return dispatch_gemm<EPI_CE_DLOGITS>(Y, W + (int64_t)v0 * ldw, ldy, ldw, 0, 0, p, (cudaStream_t)stream);
}

namespace egom2p {
__global__ void ce_finalize_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ tl,
                                   int R, int n_tiles, float* __restrict__ lse, float* __restrict__ loss_sum) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (r < R) {
    float m = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, pm[(int64_t)t * R + r]);
    float s = 0.f;
    for (int t = 0; t < n_tiles; ++t) s += ps[(int64_t)t * R + r] * __expf(pm[(int64_t)t * R + r] - m);
    const float l = m + logf(s);
    lse[r] = l;
    contrib = l - tl[r];
  }
  contrib = warp_sum(contrib);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = contrib;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, t);
  }
}
}  // namespace egom2p

extern "C" int egom2p_ce_finalize(const float* part_max, const float* part_sum, const float* tgt_logit, int32_t R,
                                  int32_t n_tiles, float* lse, float* loss_sum, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(part_max && part_sum && tgt_logit && lse && R > 0 && n_tiles > 0, "ce_finalize: bad argument");
  ce_finalize_kernel<<<(R + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part_max, part_sum, tgt_logit, R, n_tiles, lse, loss_sum);
  return check_launch("ce_finalize");
}
