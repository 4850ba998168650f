// Index plan: the reference's cat -> argsort(mask + arange*1e-6) -> gather chain (egom2p_model.py:251-283,344-481)
// collapses to a stable partition of the concatenated mask, done here with one block-wide prefix sum per sample.
// No embedding rows are touched: the plan only emits (modality, position) per kept slot, pad flags, modality ids,
// target ids and -- for the decoder -- the contiguous key range each row may attend.
#include "common.cuh"

namespace egom2p {

constexpr int kPlanThreads = 1024;

struct PlanParams {
  egom2p_plan_desc d;
  int L;
  int off[EGOM2P_MAX_MODS + 1];
  int32_t* keep_idx;
  int32_t* keep_mod;
  int32_t* keep_pos;
  uint8_t* pad;
  int16_t* mod_mask;
  int32_t* n_valid;
  int64_t* target_ids;
  int32_t* key_lo;
  int32_t* key_hi;
};

// Block-wide exclusive scan of one int per thread; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;  // exclusive warp offsets
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  int res = inc - v + s_warp[warp];
  *total = s_warp[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kPlanThreads) plan_kernel(PlanParams p) {
  __shared__ int s_warp[33];
  __shared__ int s_valid[EGOM2P_MAX_MODS];
  __shared__ int s_seg_lo[EGOM2P_MAX_MODS], s_seg_hi[EGOM2P_MAX_MODS];
  __shared__ int s_nvalid;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nm = p.d.n_mods, budget = p.d.budget;
  if (tid < EGOM2P_MAX_MODS) s_valid[tid] = 0;
  __syncthreads();

  auto locate = [&](int idx, int& m, int& pos) {
    m = 0;
    while (m + 1 < nm && idx >= p.off[m + 1]) ++m;
    pos = idx - p.off[m];
  };

  // pass 1: valid count per modality
  for (int idx = tid; idx < p.L; idx += kPlanThreads) {
    int m, pos;
    locate(idx, m, pos);
    if (!p.d.mask[m][(size_t)b * p.d.len[m] + pos]) atomicAdd(&s_valid[m], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int tot = 0, kept = 0;
    for (int m = 0; m < nm; ++m) {
      tot += s_valid[m];
      int k = min(s_valid[m], budget - kept);
      s_seg_lo[m] = kept;
      kept += k;
      s_seg_hi[m] = kept;
    }
    s_nvalid = tot;
    p.n_valid[b] = min(tot, budget);
  }
  __syncthreads();
  const int nvalid = s_nvalid;

  // pass 2: stable partition -> slots
  int running = 0;
  for (int base = 0; base < p.L; base += kPlanThreads) {
    const int idx = base + tid;
    int m = 0, pos = 0, v = 0;
    if (idx < p.L) {
      locate(idx, m, pos);
      v = p.d.mask[m][(size_t)b * p.d.len[m] + pos] ? 0 : 1;
    }
    int total;
    const int rank_v = running + block_excl_scan(v, s_warp, &total);
    running += total;
    if (idx < p.L) {
      const int slot = v ? rank_v : nvalid + (idx - rank_v);
      if (slot < budget) {
        const size_t o = (size_t)b * budget + slot;
        p.keep_idx[o] = idx;
        p.keep_mod[o] = m;
        p.keep_pos[o] = pos;
        p.pad[o] = v ? 0 : 1;
        p.mod_mask[o] = v ? (int16_t)p.d.mod_id[m] : (int16_t)-1;
        if (p.d.is_decoder) {
          const size_t src = (size_t)b * p.d.len[m] + pos;
          p.target_ids[o] = v ? p.d.ids[m][src] : 0;
          p.key_hi[o] = p.d.attn_cnt[m] ? p.d.attn_cnt[m][src] : 0;  // staged; replaced by the range below
        }
      }
    }
  }
  if (!p.d.is_decoder) return;
  __syncthreads();

  // pass 3: cumsum of the gathered attention counts -> per-row key range (adapt_decoder_attention_mask)
  running = 0;
  for (int base = 0; base < budget; base += kPlanThreads) {
    const int s = base + tid;
    const size_t o = (size_t)b * budget + s;
    const int c = (s < budget) ? p.key_hi[o] : 0;
    int total;
    const int cum = running + block_excl_scan(c, s_warp, &total) + c;
    running += total;
    if (s < budget) {
      int lo = 0, hi = p.d.causal ? s + 1 : min(cum, budget);
      if (p.d.sep) {
        const int m = p.keep_mod[o];
        lo = max(lo, s_seg_lo[m]);
        hi = min(hi, s_seg_hi[m]);
      }
      if (hi < lo) hi = lo;
      p.key_lo[o] = lo;
      p.key_hi[o] = hi;
    }
  }
}

// Rows of each modality in the compacted decoder sequence: CTA m scans mod_mask (B * budget int16) for mod_id[m] and writes
// the matching flat row indices, ascending, into rows[m * cap ..] and their number into counts[m] -- the boolean
// row-select y[decoder_mod_mask == id] (egom2p_model.py:633) for all modalities in one launch and without a
// device -> host round trip per modality.
struct RowsParams {
  const int16_t* mod_mask;
  int64_t total;
  int64_t cap;
  int32_t mod_id[EGOM2P_MAX_MODS];
  int64_t* rows;
  int32_t* counts;
};
__global__ void __launch_bounds__(kPlanThreads) plan_rows_kernel(RowsParams p) {
  __shared__ int s_warp[33];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int16_t id = (int16_t)p.mod_id[m];
  int64_t* out = p.rows + (int64_t)m * p.cap;
  int running = 0;
  for (int64_t base = 0; base < p.total; base += kPlanThreads) {
    const int64_t i = base + tid;
    const int v = (i < p.total && p.mod_mask[i] == id) ? 1 : 0;
    int total;
    const int pos = running + block_excl_scan(v, s_warp, &total);
    running += total;
    if (v) out[pos] = i;
  }
  if (tid == 0) p.counts[m] = running;
}

}  // namespace egom2p

extern "C" int egom2p_plan_rows(const int16_t* mod_mask, int64_t total, const int32_t* mod_ids, int32_t n_mods, int64_t cap,
                                int64_t* rows, int32_t* counts, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(mod_mask && mod_ids && rows && counts && total > 0 && cap >= total, "plan_rows: bad argument");
  EGO_REQUIRE(n_mods >= 1 && n_mods <= EGOM2P_MAX_MODS, "plan_rows: n_mods out of range");
  RowsParams p;
  p.mod_mask = mod_mask; p.total = total; p.cap = cap; p.rows = rows; p.counts = counts;
  for (int m = 0; m < n_mods; ++m) p.mod_id[m] = mod_ids[m];
  plan_rows_kernel<<<n_mods, kPlanThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("plan_rows");
}

extern "C" int egom2p_index_plan(const egom2p_plan_desc* desc, int32_t* keep_idx, int32_t* keep_mod, int32_t* keep_pos,
                                 uint8_t* pad, int16_t* mod_mask, int32_t* n_valid, int64_t* target_ids,
                                 int32_t* key_lo, int32_t* key_hi, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(desc && desc->n_mods >= 1 && desc->n_mods <= EGOM2P_MAX_MODS, "index_plan: n_mods out of range");
  EGO_REQUIRE(desc->batch >= 1 && desc->budget >= 1, "index_plan: empty batch or budget");
  EGO_REQUIRE(keep_idx && keep_mod && keep_pos && pad && mod_mask && n_valid, "index_plan: null output");
  PlanParams p;
  p.d = *desc;
  p.off[0] = 0;
  for (int m = 0; m < desc->n_mods; ++m) {
    EGO_REQUIRE(desc->mask[m] && desc->len[m] > 0, "index_plan: modality %d has no mask", m);
    if (desc->is_decoder) EGO_REQUIRE(desc->ids[m] != nullptr, "index_plan: decoder plan needs ids");
    p.off[m + 1] = p.off[m] + desc->len[m];
  }
  p.L = p.off[desc->n_mods];
  EGO_REQUIRE(desc->budget <= p.L, "index_plan: budget %d exceeds the %d available positions", desc->budget, p.L);
  if (desc->is_decoder) EGO_REQUIRE(target_ids && key_lo && key_hi, "index_plan: decoder outputs missing");
  p.keep_idx = keep_idx; p.keep_mod = keep_mod; p.keep_pos = keep_pos; p.pad = pad; p.mod_mask = mod_mask;
  p.n_valid = n_valid; p.target_ids = target_ids; p.key_lo = key_lo; p.key_hi = key_hi;
  plan_kernel<<<desc->batch, kPlanThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("index_plan");
}
