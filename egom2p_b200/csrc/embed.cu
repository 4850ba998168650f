// Fused per-modality token-embedding + positional/modality-embedding + masked gather (north-star kernel 1).
// One warp per kept slot; rows are 16-byte vectorised and each warp instruction touches 512 contiguous bytes.
// Arithmetic order matches the reference exactly: emb = pos + mod; x0 = tok + emb; pads = 0.
#include "common.cuh"

namespace egom2p {

struct EmbedParams {
  egom2p_embed_desc d;
  const float* mask_token;
  const int32_t* keep_mod;
  const int32_t* keep_pos;
  const uint8_t* pad;
  const int32_t* row_batch;   // sample of each row, or NULL: row / budget (rows are (B, budget) slots)
  int64_t rows;
  int32_t budget;
  float* x0;
  float* emb;
};

__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedParams p) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const int lane = threadIdx.x & 31;
  const int D4 = p.d.dim >> 2;
  float4* xo = reinterpret_cast<float4*>(p.x0 + row * p.d.dim);
  float4* eo = p.emb ? reinterpret_cast<float4*>(p.emb + row * p.d.dim) : nullptr;
  if (p.pad[row]) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = lane; i < D4; i += 32) {
      xo[i] = z;
      if (eo) eo[i] = z;
    }
    return;
  }
  const int m = p.keep_mod[row], pos = p.keep_pos[row];
  const int64_t b = p.row_batch ? p.row_batch[row] : row / p.budget;
  const float4* pe = reinterpret_cast<const float4*>(p.d.pos_emb[m] + (size_t)pos * p.d.dim);
  const float4* me = reinterpret_cast<const float4*>(p.d.mod_emb[m]);
  const float4* te;
  if (p.mask_token) {
    te = reinterpret_cast<const float4*>(p.mask_token);
  } else {
    int64_t id = p.d.ids[m][b * p.d.len[m] + pos];
    id = id < 0 ? 0 : (id >= p.d.vocab[m] ? p.d.vocab[m] - 1 : id);
    te = reinterpret_cast<const float4*>(p.d.token_emb[m] + (size_t)id * p.d.dim);
  }
  for (int i = lane; i < D4; i += 32) {
    const float4 a = __ldg(pe + i), c = __ldg(me + i), t = __ldg(te + i);
    float4 e, x;
    e.x = a.x + c.x; e.y = a.y + c.y; e.z = a.z + c.z; e.w = a.w + c.w;
    if (p.mask_token) {  // reference: (zeros + mask_token) + emb
      x.x = (0.f + t.x) + e.x; x.y = (0.f + t.y) + e.y; x.z = (0.f + t.z) + e.z; x.w = (0.f + t.w) + e.w;
    } else {
      x.x = t.x + e.x; x.y = t.y + e.y; x.z = t.z + e.z; x.w = t.w + e.w;
    }
    xo[i] = x;
    if (eo) eo[i] = e;
  }
}

struct EmbedBwdParams {
  egom2p_embed_desc d;
  const float* dx0;
  const float* demb;
  const int32_t* keep_mod;
  const int32_t* keep_pos;
  const uint8_t* pad;
  const int32_t* row_batch;
  int64_t rows;
  int32_t budget;
  float* d_token_emb[EGOM2P_MAX_MODS];
  float* d_mod_emb[EGOM2P_MAX_MODS];
  float* d_mask_token;
  int rows_per_cta;
};

// Each CTA walks `rows_per_cta` consecutive slots: token-table rows get an atomic scatter-add (row-sparse grad),
// modality / mask-token sums are accumulated per thread (thread owns fixed columns) and flushed once per modality
// change, so the dense (n_mods, dim) outputs see O(#CTAs * n_mods) atomics per column.
__global__ void __launch_bounds__(256) embed_bwd_kernel(EmbedBwdParams p) {
  const int D = p.d.dim;
  const int64_t r0 = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t r1 = min(r0 + (int64_t)p.rows_per_cta, p.rows);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc_mod = 0.f, acc_tok = 0.f;
    int cur = -1;
    for (int64_t r = r0; r < r1; ++r) {
      if (p.pad[r]) continue;
      const int m = p.keep_mod[r];
      if (m != cur) {
        if (cur >= 0 && p.d_mod_emb[cur]) atomicAdd(p.d_mod_emb[cur] + c, acc_mod);
        acc_mod = 0.f;
        cur = m;
      }
      const float g = p.dx0[r * D + c];
      acc_mod += g + (p.demb ? p.demb[r * D + c] : 0.f);
      if (p.d_mask_token) {
        acc_tok += g;
      } else if (p.d_token_emb[m]) {
        const int64_t b = p.row_batch ? p.row_batch[r] : r / p.budget;
        int64_t id = p.d.ids[m][b * p.d.len[m] + p.keep_pos[r]];
        id = id < 0 ? 0 : (id >= p.d.vocab[m] ? p.d.vocab[m] - 1 : id);
        atomicAdd(p.d_token_emb[m] + (size_t)id * D + c, g);
      }
    }
    if (cur >= 0 && p.d_mod_emb[cur]) atomicAdd(p.d_mod_emb[cur] + c, acc_mod);
    if (p.d_mask_token) atomicAdd(p.d_mask_token + c, acc_tok);
  }
}

}  // namespace egom2p

extern "C" int egom2p_embed_gather_fwd(const egom2p_embed_desc* desc, const float* mask_token, const int32_t* keep_mod,
                                       const int32_t* keep_pos, const uint8_t* pad, const int32_t* row_batch, int64_t rows,
                                       int32_t budget, float* x0, float* emb, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(desc && desc->n_mods >= 1 && desc->n_mods <= EGOM2P_MAX_MODS, "embed_gather_fwd: n_mods out of range");
  EGO_REQUIRE(desc->dim > 0 && desc->dim % 4 == 0, "embed_gather_fwd: dim must be a multiple of 4");
  EGO_REQUIRE(keep_mod && keep_pos && pad && x0 && rows > 0 && budget > 0, "embed_gather_fwd: null argument");
  for (int m = 0; m < desc->n_mods; ++m) {
    EGO_REQUIRE(desc->pos_emb[m] && desc->mod_emb[m], "embed_gather_fwd: modality %d tables missing", m);
    if (!mask_token) EGO_REQUIRE(desc->ids[m] && desc->token_emb[m], "embed_gather_fwd: modality %d ids/table missing", m);
  }
  EmbedParams p{*desc, mask_token, keep_mod, keep_pos, pad, row_batch, rows, budget, x0, emb};
  embed_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("embed_gather_fwd");
}

extern "C" int egom2p_embed_gather_bwd(const egom2p_embed_desc* desc, const float* dx0, const float* demb,
                                       const int32_t* keep_mod, const int32_t* keep_pos, const uint8_t* pad,
                                       const int32_t* row_batch, int64_t rows, int32_t budget, float* const* d_token_emb,
                                       float* const* d_mod_emb, float* d_mask_token, void* stream) {
  using namespace egom2p;
  EGO_REQUIRE(desc && desc->n_mods >= 1 && desc->n_mods <= EGOM2P_MAX_MODS, "embed_gather_bwd: n_mods out of range");
  EGO_REQUIRE(dx0 && keep_mod && keep_pos && pad && rows > 0 && budget > 0, "embed_gather_bwd: null argument");
  EmbedBwdParams p;
  p.d = *desc; p.dx0 = dx0; p.demb = demb; p.keep_mod = keep_mod; p.keep_pos = keep_pos; p.pad = pad; p.row_batch = row_batch;
  p.rows = rows; p.budget = budget; p.d_mask_token = d_mask_token;
  for (int m = 0; m < EGOM2P_MAX_MODS; ++m) {
    p.d_token_emb[m] = (d_token_emb && m < desc->n_mods) ? d_token_emb[m] : nullptr;
    p.d_mod_emb[m] = (d_mod_emb && m < desc->n_mods) ? d_mod_emb[m] : nullptr;
    if (p.d_token_emb[m]) EGO_REQUIRE(desc->ids[m] != nullptr, "embed_gather_bwd: ids missing for modality %d", m);
  }
  p.rows_per_cta = 32;
  const unsigned grid = (unsigned)((rows + p.rows_per_cta - 1) / p.rows_per_cta);
  embed_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("embed_gather_bwd");
}
