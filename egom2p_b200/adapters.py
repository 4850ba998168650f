"""Per-modality encoder / decoder embedding adapters with the reference's names, constructor arguments, attributes
and state_dict layout (reference: egom2p/models/encoder_embeddings.py:124-301, decoder_embeddings.py:271-500).

In the fused training path the model reads the adapters' tables directly (token_emb.weight, pos_emb, mod_emb,
to_logits.weight) -- the reference's own adapter objects work just as well there (duck-typed). The forward methods
below serve the GenerationSampler call surface (`encoder_embeddings[mod](d)`, `decoder_embeddings[mod].forward_embed(d)`,
`.forward_logits(y)`, `.token_emb(ids)`) and run on the CUDA kernels (no eager fallback).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple, Union

import torch
from torch import nn

from . import ops
from .posemb import build_1d_sincos_posemb, build_3d_sincos_posemb


def _pair(t):
    return t if isinstance(t, tuple) else (t, t)


class _TokenEmbedding(nn.Embedding):
    """nn.Embedding container whose lookup runs on the fused gather kernel."""

    def forward(self, ids: torch.Tensor) -> torch.Tensor:  # used by the sampler for autoregressive paths
        return _gather_rows(self.weight, ids)


def _gather_rows(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """out[..., :] = table[ids] through the embed-gather kernel (identity plan, zero pos/mod tables)."""
    shape = ids.shape
    flat = ids.reshape(1, -1).to(torch.int64).contiguous()
    n, dim = flat.shape[1], table.shape[1]
    plan = ops.Plan()
    plan.B, plan.budget = 1, n
    dev = table.device
    plan.keep_mod = torch.zeros(1, n, dtype=torch.int32, device=dev)
    plan.keep_pos = torch.arange(n, dtype=torch.int32, device=dev).reshape(1, n)
    plan.pad = torch.zeros(1, n, dtype=torch.bool, device=dev)
    zeros_pos = torch.zeros(n, dim, dtype=torch.float32, device=dev)
    zeros_mod = torch.zeros(dim, dtype=torch.float32, device=dev)
    x, _ = ops.embed_gather_fwd(plan, dim, [n], [table.shape[0]], [flat], [table.detach()], [zeros_pos], [zeros_mod],
                                want_emb=False)
    return x.reshape(*shape, dim)


class _AdapterBase(nn.Module):
    def _init_common(self, n_pos_emb: torch.Tensor, init_std: float):
        if self.sincos_pos_emb:
            self.register_buffer("pos_emb", n_pos_emb)
        else:
            self.pos_emb = nn.Parameter(torch.zeros_like(n_pos_emb))
            nn.init.normal_(self.pos_emb, std=init_std)
        self.mod_emb = nn.Parameter(torch.zeros(1, 1, self.dim_tokens))
        nn.init.normal_(self.mod_emb, std=init_std)
        self.token_emb = _TokenEmbedding(num_embeddings=self.vocab_size, embedding_dim=self.dim_tokens)

    @torch.jit.ignore
    def no_weight_decay(self):
        return set()

    def _embed(self, d: Dict[str, torch.Tensor], with_ids: bool) -> Dict[str, torch.Tensor]:
        ids = d["tensor"]
        B = ids.shape[0]
        ids = ids.reshape(B, -1)
        d["x"] = self.token_emb(ids)
        emb, _ = ops.add_f32(self.pos_emb.detach().contiguous(), self.mod_emb.detach().expand_as(self.pos_emb).contiguous())
        d["emb"] = emb.expand(B, -1, -1)
        if with_ids:
            d["ids"] = ids
        assert d["x"].shape[1] == d["emb"].shape[1]
        return d


class GazeCamTokenEncoderEmbedding(_AdapterBase):
    """30-token camera-trajectory / gaze streams, vocab 256, 1-D sin-cos positions (encoder_embeddings.py:124-210)."""

    def __init__(self, vocab_size: int = 256, dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, **kwargs):
        super().__init__()
        self.vocab_size, self.dim_tokens, self.sincos_pos_emb = vocab_size, dim_tokens, sincos_pos_emb
        if dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        self._init_common(build_1d_sincos_posemb(30, embed_dim=dim_tokens), init_std)

    def forward(self, d):
        return self._embed(d, with_ids=False)


class VideoTokenEncoderEmbedding(_AdapterBase):
    """Cosmos DV4x8x8 video tokens (5 x 32 x 32 for a 16x256x256 clip), vocab 64000, 3-D sin-cos positions
    (encoder_embeddings.py:212-301)."""

    def __init__(self, vocab_size: int = 64000, patch_size: Union[int, Tuple[int, int, int]] = (4, 8, 8),
                 dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, image_size: Union[int, Tuple[int]] = 256,
                 **kwargs):
        super().__init__()
        self.vocab_size, self.patch_size, self.dim_tokens = vocab_size, patch_size, dim_tokens
        self.sincos_pos_emb, self.image_size = sincos_pos_emb, _pair(image_size)
        if dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        h, w = self.image_size[0] // self.patch_size[1], self.image_size[1] // self.patch_size[2]
        self._init_common(build_3d_sincos_posemb(t=5, h=h, w=w, embed_dim=dim_tokens), init_std)

    def forward(self, d):
        return self._embed(d, with_ids=False)


class _DecoderMixin:
    def _init_head(self):
        self.to_logits = _HeadLinear(self.dim_tokens, self.vocab_size, bias=False)
        if self.share_embedding:
            self.to_logits.weight = self.token_emb.weight

    def forward_embed(self, d):
        return self._embed(d, with_ids=True)

    def forward_logits(self, x: torch.Tensor) -> torch.Tensor:
        return self.to_logits(x)


class _HeadLinear(nn.Linear):
    """Vocabulary projection container; standalone calls (sampler) run the tcgen05 GEMM and return fp32 logits."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if x2.shape[0] == 0:
            return x.new_zeros(*shape[:-1], self.weight.shape[0])
        xb = x2 if x2.dtype == torch.bfloat16 else ops.cast_bf16(x2.float().contiguous())
        wb = ops.cached_bf16(self.weight)   # re-cast only when the master changed (ops.cached_bf16)
        return ops.linear_fwd(xb, wb, out_dtype=torch.float32).reshape(*shape[:-1], -1)


class GazeCamTokenDecoderEmbedding(_AdapterBase, _DecoderMixin):
    """decoder_embeddings.py:271-383."""

    def __init__(self, vocab_size: int = 256, patch_size: Union[int, Tuple[int, int]] = 2, dim_tokens: Optional[int] = None,
                 sincos_pos_emb: bool = True, share_embedding: bool = True, **kwargs):
        super().__init__()
        self.vocab_size, self.patch_size, self.dim_tokens = vocab_size, patch_size, dim_tokens
        self.sincos_pos_emb, self.share_embedding = sincos_pos_emb, share_embedding
        if dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        self._init_common(build_1d_sincos_posemb(30, embed_dim=dim_tokens), init_std)
        self._init_head()


class VideoTokenDecoderEmbedding(_AdapterBase, _DecoderMixin):
    """decoder_embeddings.py:385-500."""

    def __init__(self, vocab_size: int = 64000, patch_size: Union[int, Tuple[int, int, int]] = (4, 8, 8),
                 dim_tokens: Optional[int] = None, sincos_pos_emb: bool = True, image_size: Union[int, Tuple[int]] = 256,
                 share_embedding: bool = True, **kwargs):
        super().__init__()
        self.vocab_size, self.patch_size, self.dim_tokens = vocab_size, patch_size, dim_tokens
        self.sincos_pos_emb, self.image_size, self.share_embedding = sincos_pos_emb, _pair(image_size), share_embedding
        if dim_tokens is not None:
            self.init(dim_tokens=dim_tokens)

    def init(self, dim_tokens: int = 768, init_std=0.02):
        self.dim_tokens = dim_tokens
        h, w = self.image_size[0] // self.patch_size[1], self.image_size[1] // self.patch_size[2]
        self._init_common(build_3d_sincos_posemb(t=5, h=h, w=w, embed_dim=dim_tokens), init_std)
        self._init_head()
