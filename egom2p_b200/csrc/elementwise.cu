// HBM-bound elementwise kernels on the path: SwiGLU gate (egom2p_utils.py:167-169), casts, residual add,
// fused AdamW (torch.optim.AdamW semantics) and sum-of-squares for the global grad-norm clip.
// All are grid-stride, 16-byte vectorised, grid = 148 SMs x 8 CTAs.
#include "common.cuh"

namespace egom2p {

constexpr int kEwThreads = 256;
static inline unsigned ew_grid(int64_t nvec) {
  int64_t g = (nvec + kEwThreads - 1) / kEwThreads;
  const int64_t cap = 148 * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  r.x = pack_bf16(f[0], f[1]); r.y = pack_bf16(f[2], f[3]); r.z = pack_bf16(f[4], f[5]); r.w = pack_bf16(f[6], f[7]);
  return r;
}

__global__ void __launch_bounds__(kEwThreads) swiglu_fwd_kernel(const uint16_t* __restrict__ ab, int64_t rows, int hidden,
                                                                uint16_t* __restrict__ g) {
  const int H8 = hidden >> 3;
  const int64_t n = rows * H8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H8;
    const int c = (int)(i - r * H8);
    const uint4 ra = reinterpret_cast<const uint4*>(ab + r * 2 * hidden)[c];
    const uint4 rb = reinterpret_cast<const uint4*>(ab + r * 2 * hidden + hidden)[c];
    float a[8], b[8], o[8];
    unpack8(ra, a);
    unpack8(rb, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      o[k] = a[k] / (1.f + __expf(-a[k])) * b[k];
    }
    reinterpret_cast<uint4*>(g + r * hidden)[c] = pack8(o);
  }
}

__global__ void __launch_bounds__(kEwThreads) swiglu_bwd_kernel(const uint16_t* __restrict__ ab, const uint16_t* __restrict__ dg,
                                                                int64_t rows, int hidden, uint16_t* __restrict__ dab) {
  const int H8 = hidden >> 3;
  const int64_t n = rows * H8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / H8;
    const int c = (int)(i - r * H8);
    const uint4 ra = reinterpret_cast<const uint4*>(ab + r * 2 * hidden)[c];
    const uint4 rb = reinterpret_cast<const uint4*>(ab + r * 2 * hidden + hidden)[c];
    const uint4 rg = reinterpret_cast<const uint4*>(dg + r * hidden)[c];
    float a[8], b[8], d[8], da[8], db[8];
    unpack8(ra, a);
    unpack8(rb, b);
    unpack8(rg, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float sg = 1.f / (1.f + __expf(-a[k]));
      const float s = a[k] * sg;
      db[k] = d[k] * s;
      da[k] = d[k] * b[k] * (sg * (1.f + a[k] * (1.f - sg)));
    }
    reinterpret_cast<uint4*>(dab + r * 2 * hidden)[c] = pack8(da);
    reinterpret_cast<uint4*>(dab + r * 2 * hidden + hidden)[c] = pack8(db);
  }
}

__global__ void __launch_bounds__(kEwThreads) cast_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t n) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 pk;
    pk.x = pack_bf16(v.x, v.y);
    pk.y = pack_bf16(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = pk;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    __nv_bfloat16 h = __float2bfloat16(src[i]);
    dst[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

// One launch that re-casts a whole list of fp32 master weights into their bf16 GEMM operands (the per-step refresh after the
// optimizer wrote the masters). A block = one chunk of rows of one tensor, found by bisection over the items' first chunks;
// group > 0 writes the rows into the interleaved [group x fc1 | group x fc3] layout the SwiGLU-epilogue GEMM consumes.
struct CastItem {
  const float* src;
  uint16_t* dst;
  int64_t rows;
  int64_t first_chunk;
  int32_t cols, dst_ld, group, slot, rows_per_chunk, pad_;
};
static_assert(sizeof(CastItem) == 56, "CastItem layout is part of the C ABI (egom2p_cast_item)");
__global__ void __launch_bounds__(kEwThreads) cast_multi_kernel(const CastItem* __restrict__ items, int n_items) {
  __shared__ CastItem it;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {  // last item with first_chunk <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].first_chunk <= (int64_t)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    it = items[lo];
  }
  __syncthreads();
  const int64_t r0 = ((int64_t)blockIdx.x - it.first_chunk) * it.rows_per_chunk;
  const int nr = (int)min((int64_t)it.rows_per_chunk, it.rows - r0);
  auto dst_row = [&](int64_t r) -> int64_t {
    return it.group > 0 ? (r / it.group) * 2 * it.group + (int64_t)it.slot * it.group + r % it.group : r;
  };
  if ((it.cols & 3) == 0 && (it.dst_ld & 3) == 0) {
    const int c4 = it.cols >> 2;
    for (int i = threadIdx.x; i < nr * c4; i += kEwThreads) {
      const int rr = i / c4, cc = i - rr * c4;
      const float4 v = reinterpret_cast<const float4*>(it.src + (r0 + rr) * it.cols)[cc];
      uint2 pk;
      pk.x = pack_bf16(v.x, v.y);
      pk.y = pack_bf16(v.z, v.w);
      reinterpret_cast<uint2*>(it.dst + dst_row(r0 + rr) * it.dst_ld)[cc] = pk;
    }
  } else {
    for (int i = threadIdx.x; i < nr * it.cols; i += kEwThreads) {
      const int rr = i / it.cols, cc = i - rr * it.cols;
      const __nv_bfloat16 h = __float2bfloat16(it.src[(r0 + rr) * it.cols + cc]);
      it.dst[dst_row(r0 + rr) * it.dst_ld + cc] = *reinterpret_cast<const uint16_t*>(&h);
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                         float* __restrict__ out, uint16_t* __restrict__ outb) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(a)[i], y = reinterpret_cast<const float4*>(b)[i];
    const float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    if (out) reinterpret_cast<float4*>(out)[i] = o;
    if (outb) {
      uint2 pk;
      pk.x = pack_bf16(o.x, o.y);
      pk.y = pack_bf16(o.z, o.w);
      reinterpret_cast<uint2*>(outb)[i] = pk;
    }
  }
}

__global__ void __launch_bounds__(kEwThreads) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                           float* __restrict__ v, int64_t n, float lr, float b1, float b2,
                                                           float eps, float wd, float bc1, float bc2_sqrt,
                                                           const float* __restrict__ gscale) {
  const float gs = gscale ? *gscale : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}

__global__ void __launch_bounds__(kEwThreads) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i] * x[i];
  s = warp_sum(s);
  __shared__ float sh[kEwThreads / 32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < kEwThreads / 32 ? sh[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// out[c] += sum_r x[r, c]; each CTA reduces a 64-row slab, threads own columns (coalesced), one atomic per column.
__global__ void __launch_bounds__(kEwThreads) colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, float* __restrict__ out) {
  const int64_t r0 = (int64_t)blockIdx.x * 64, r1 = min(r0 + 64, rows);
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += x[r * cols + c];
    atomicAdd(out + c, s);
  }
}

// Row gather / scatter-add used to (un)compact the rows of one modality for its vocabulary head.
__global__ void __launch_bounds__(kEwThreads) gather_rows_bf16_kernel(const uint16_t* __restrict__ src, const int64_t* __restrict__ idx,
                                                                      int64_t n, int cols8, uint16_t* __restrict__ dst) {
  const int64_t total = n * cols8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols8;
    const int c = (int)(i - r * cols8);
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[idx[r] * cols8 + c];
  }
}
__global__ void __launch_bounds__(kEwThreads) scatter_rows_f32_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                                      int64_t n, int cols4, float* __restrict__ dst) {
  const int64_t total = n * cols4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4);
    reinterpret_cast<float4*>(dst)[idx[r] * cols4 + c] = reinterpret_cast<const float4*>(src)[i];
  }
}

}  // namespace egom2p

extern "C" int egom2p_colsum_f32(const float* x, int64_t rows, int32_t cols, float* out, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(x && out && rows > 0 && cols > 0, "colsum_f32: bad argument");
  colsum_kernel<<<(unsigned)((rows + 63) / 64), kEwThreads, 0, (cudaStream_t)stream>>>(x, rows, cols, out);
  return check_launch("colsum_f32");
}
extern "C" int egom2p_gather_rows_bf16(const uint16_t* src, const int64_t* idx, int64_t n, int32_t cols, uint16_t* dst, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(src && idx && dst && n > 0 && cols > 0 && cols % 8 == 0, "gather_rows_bf16: bad argument (cols %% 8 == 0)");
  gather_rows_bf16_kernel<<<ew_grid(n * (cols / 8)), kEwThreads, 0, (cudaStream_t)stream>>>(src, idx, n, cols / 8, dst);
  return check_launch("gather_rows_bf16");
}
extern "C" int egom2p_scatter_rows_f32(const float* src, const int64_t* idx, int64_t n, int32_t cols, float* dst, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(src && idx && dst && n > 0 && cols > 0 && cols % 4 == 0, "scatter_rows_f32: bad argument (cols %% 4 == 0)");
  scatter_rows_f32_kernel<<<ew_grid(n * (cols / 4)), kEwThreads, 0, (cudaStream_t)stream>>>(src, idx, n, cols / 4, dst);
  return check_launch("scatter_rows_f32");
}

extern "C" int egom2p_swiglu_fwd(const uint16_t* ab, int64_t rows, int32_t hidden, uint16_t* g, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(ab && g && rows > 0 && hidden > 0 && hidden % 8 == 0, "swiglu_fwd: bad argument (hidden %% 8 == 0 required)");
  swiglu_fwd_kernel<<<ew_grid(rows * (hidden / 8)), kEwThreads, 0, (cudaStream_t)stream>>>(ab, rows, hidden, g);
  return check_launch("swiglu_fwd");
}
extern "C" int egom2p_swiglu_bwd(const uint16_t* ab, const uint16_t* dg, int64_t rows, int32_t hidden, uint16_t* dab, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(ab && dg && dab && rows > 0 && hidden > 0 && hidden % 8 == 0, "swiglu_bwd: bad argument");
  swiglu_bwd_kernel<<<ew_grid(rows * (hidden / 8)), kEwThreads, 0, (cudaStream_t)stream>>>(ab, dg, rows, hidden, dab);
  return check_launch("swiglu_bwd");
}
extern "C" int egom2p_cast_f32_to_bf16(const float* src, uint16_t* dst, int64_t n, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(src && dst && n > 0, "cast_f32_to_bf16: bad argument");
  EGO_REQUIRE(((uintptr_t)src & 15) == 0 && ((uintptr_t)dst & 7) == 0, "cast_f32_to_bf16: misaligned");
  cast_kernel<<<ew_grid(n / 4 + 1), kEwThreads, 0, (cudaStream_t)stream>>>(src, dst, n);
  return check_launch("cast_f32_to_bf16");
}
extern "C" int egom2p_cast_f32_to_bf16_multi(const void* items_dev, int32_t n_items, int64_t n_chunks, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(items_dev && n_items > 0 && n_chunks > 0 && n_chunks < (int64_t)INT32_MAX, "cast_f32_to_bf16_multi: bad argument");
  cast_multi_kernel<<<(unsigned)n_chunks, kEwThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const CastItem*>(items_dev), n_items);
  return check_launch("cast_f32_to_bf16_multi");
}
extern "C" int egom2p_add_f32(const float* a, const float* b, int64_t n, float* out, uint16_t* out_bf16, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(a && b && (out || out_bf16) && n > 0 && n % 4 == 0, "add_f32: bad argument (n %% 4 == 0 required)");
  add_kernel<<<ew_grid(n / 4), kEwThreads, 0, (cudaStream_t)stream>>>(a, b, n, out, out_bf16);
  return check_launch("add_f32");
}
extern "C" int egom2p_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                 const float* grad_scale, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "adamw_step: bad argument");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                   weight_decay, bc1, bc2s, grad_scale);
  return check_launch("adamw_step");
}
extern "C" int egom2p_sumsq_f32(const float* x, int64_t n, float* sumsq, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(x && sumsq && n > 0, "sumsq_f32: bad argument");
  sumsq_kernel<<<ew_grid(n), kEwThreads, 0, (cudaStream_t)stream>>>(x, n, sumsq);
  return check_launch("sumsq_f32");
}
