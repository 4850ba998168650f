"""Drop-in replacement of the reference `EgoM2P` module (egom2p/models/egom2p_model.py:57-819) whose forward /
backward run on the sm_100a kernels of libegom2p_b200.so.

Boundary kept intact: constructor signature, `forward(mod_dict, num_encoder_tokens, num_decoder_tokens, loss_type,
return_logits)`, the masking-dict layout, adapter objects (reference adapters work, duck-typed), parameter / buffer
names and shapes (strict `load_state_dict` of reference checkpoints), `forward_encoder / forward_decoder /
forward_logits / decoder_proj_context / mask_token / cat_encoder_tensors` for the GenerationSampler, freeze helpers,
`no_weight_decay`. What changed is *how* the step is computed: index plan + fused embed/gather,
range-masked flash attention, tcgen05 GEMMs with fused epilogues, vocabulary head fused with cross-entropy.
Autograd nodes are per block, so DDP's bucketed all-reduce overlaps with backward exactly as for the reference.

Scope: the swiglu / no-bias family (ego-b and its tiny/small/large siblings). Other registered variants of the
reference (GELU+bias, QK-norm) are not built here and raise NotImplementedError at construction.
"""
from __future__ import annotations

import math
import random
from functools import partial
from typing import Any, Dict, List, Optional, Union

import torch
from torch import nn

from . import ops

bf16, f32 = torch.bfloat16, torch.float32


# =============================================================================================== parameter containers
class LayerNorm(nn.Module):
    """Parameter container mirroring egom2p_utils.LayerNorm (weight + zero `bias` buffer when bias=False)."""

    def __init__(self, normalized_shape: int, eps=1e-5, bias=True):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        if bias:
            raise NotImplementedError("egom2p_b200 builds the no-bias LayerNorm family only")
        self.register_buffer("bias", torch.zeros(normalized_shape))
        self.normalized_shape = (normalized_shape,)

    def forward(self, x):
        shape = x.shape
        _, y, _, _ = ops.layernorm_fwd(x.reshape(-1, shape[-1]).float().contiguous(), self.weight.detach(), self.eps,
                                       out_bf16=False, out_f32=True, save_stats=False)
        return y.reshape(shape)


class _Linear(nn.Linear):
    """nn.Linear container. Standalone calls (the sampler uses `decoder_proj_context(x)`) run the tcgen05 GEMM."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        xb = x2 if x2.dtype == bf16 else ops.cast_bf16(x2.float().contiguous())
        y = ops.linear_fwd(xb, ops.cached_bf16(self.weight), bias=None if self.bias is None else self.bias.detach(), out_dtype=f32)
        return y.reshape(*shape[:-1], -1)


class GatedMlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        hidden = int(2 * hidden / 3)
        self.fc1 = _Linear(dim, hidden, bias=False)
        self.fc2 = _Linear(hidden, dim, bias=False)
        self.fc3 = _Linear(dim, hidden, bias=False)


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = _Linear(dim, dim * 3, bias=False)
        self.proj = _Linear(dim, dim, bias=False)


class CrossAttention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = _Linear(dim, dim, bias=False)
        self.kv = _Linear(dim, dim * 2, bias=False)
        self.proj = _Linear(dim, dim, bias=False)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = GatedMlp(dim, int(dim * mlp_ratio))


class DecoderBlock(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.self_attn = Attention(dim, num_heads)
        self.cross_attn = CrossAttention(dim, num_heads)
        self.query_norm = norm_layer(dim)
        self.context_norm = norm_layer(dim)
        self.norm2 = norm_layer(dim)
        self.mlp = GatedMlp(dim, int(dim * mlp_ratio))


# =============================================================================================== kernels glue
class _Geom:
    """Shapes + key-range tables of one forward (device tensors, no host syncs)."""
    __slots__ = ("B", "N", "M", "H", "D", "enc_lo", "enc_hi", "dec_lo", "dec_hi", "x_lo", "x_hi", "eps", "m_enc", "m_dec", "m_x",
                 "ctx_users", "dctx_acc", "epoch_ref", "p_enc", "p_dec", "p_x")

    def build_meta(self, dev, encoder: bool = True):
        """Range metadata per attention kind, once per forward (shared by all layers, heads, fwd and bwd)."""
        if self.N > 0 and encoder:
            self.m_enc = ops.attn_ranges(self.B, self.N, self.N, self.enc_lo, self.enc_hi, device=dev)
        if self.M > 0:
            self.m_dec = ops.attn_ranges(self.B, self.M, self.M, self.dec_lo, self.dec_hi, device=dev)
            self.m_x = ops.attn_ranges(self.B, self.M, self.N, self.x_lo, self.x_hi, device=dev)
        self.p_enc = self.p_dec = self.p_x = None
        if ops.KernelTimer.active is not None:   # mask-aware (query, key) pair counts for the timing breakdown (device scalars)
            cnt = lambda lo, hi: None if lo is None else (hi.long() - lo.long()).clamp(min=0).sum()
            self.p_enc, self.p_dec, self.p_x = cnt(self.enc_lo, self.enc_hi), cnt(self.dec_lo, self.dec_hi), cnt(self.x_lo, self.x_hi)


def _check_epoch(ref, seen):
    """The bf16 operand buffers are re-cast IN PLACE after every optimizer step (EgoM2P._refresh_operands); a backward that
    runs after such a refresh of a forward made before it would silently use the new weights in its dgrad GEMMs. The
    reference raises a saved-tensor version error in that situation; so does this."""
    if ref[0] != seen:
        raise RuntimeError("egom2p_b200: the bf16 weight operands saved by this forward were refreshed (weights changed and "
                           "another forward ran) before its backward; run backward before the next forward of updated weights")


def _zeros(n, dev):
    return torch.zeros(n, dtype=f32, device=dev)


# Bumped by every backward of the block / context / head functions below. The bf16 operand cache of EgoM2P is keyed on it:
# weights whose gradients were just produced are about to be rewritten by an optimizer, and the fused CUDA optimizers
# (torch._fused_adamw_) do not bump Tensor._version, so the version counter alone would leave the cache stale.
_GRAD_GEN = ops.GRAD_GEN


def _split_w13_grad(dw13, F):
    """dw13 rows are interleaved [32 x fc1 | 32 x fc3] groups (see EgoM2P._bf16_w13): back to (fc1.grad, fc3.grad)."""
    Fp, D = dw13.shape[0] // 2, dw13.shape[1]
    v = dw13.view(Fp // 32, 2, 32, D)
    return v[:, 0].reshape(Fp, D)[:F], v[:, 1].reshape(Fp, D)[:F]


def _with_bf16(dx, dxb):
    """Hands the bf16 copy of a residual-stream gradient (written by the LayerNorm backward that produced it, for 1/3 of the
    traffic of a separate cast pass) to the next backward node: autograd passes the same tensor object on, attributes included."""
    dx._egom2p_bf16 = (dxb, dx._version, dx.data_ptr())
    return dx


def _bf16_of(dx):
    hit = getattr(dx, "_egom2p_bf16", None)
    if hit is not None:
        dxb, ver, ptr = hit
        # an in-place accumulation into dx (a second consumer of the residual stream, a hook) bumps its version: the copy is
        # stale then and is rebuilt from the fp32 gradient
        if ver == dx._version and ptr == dx.data_ptr() and dxb.shape == dx.shape and dxb.device == dx.device and dxb.dtype == bf16:
            return dxb
    return ops.cast_bf16(dx)


def _mlp_fwd(x1, n2w, w13, w2, eps):
    h2, _, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, eps)
    ab, g = ops.gemm_swiglu_fwd(h2, w13)   # fc1|fc3 GEMM with the SwiGLU gate in its epilogue
    x2 = ops.linear_fwd(g, w2, addend=x1, out_dtype=f32)
    return x2, (mean2, rstd2, h2, ab, g)


def _mlp_bwd(dx2, x1, n2w, w13, w2, saved):
    """Returns dx1 (incl. the residual path), its bf16 copy, dn2w, dw13, dw2."""
    mean2, rstd2, h2, ab, g = saved
    dx2b = _bf16_of(dx2)
    dw2 = ops.linear_wgrad(dx2b, g)
    dab = ops.gemm_swiglu_bwd(dx2b, w2, ab)   # fc2 dgrad with the SwiGLU derivative in its epilogue
    dw13 = ops.linear_wgrad(dab, h2)
    dh2 = ops.linear_dgrad(dab, w13)
    dn2w = _zeros(n2w.numel(), dx2.device)
    dx1, dx1b = ops.layernorm_bwd(dh2, x1, n2w, mean2, rstd2, dx_in=dx2, d_weight=dn2w, want_bf16=True)
    return dx1, dx1b, dn2w, dw13, dw2


def _self_attn_fwd(x, n1w, wqkv, wproj, B, L, H, meta, eps, pairs=None):
    D = x.shape[-1]
    h1, _, mean1, rstd1 = ops.layernorm_fwd(x, n1w, eps)
    qkv = ops.linear_fwd(h1, wqkv)
    o, lse = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, L, L, meta=meta, pairs=pairs)
    x1 = ops.linear_fwd(o, wproj, addend=x, out_dtype=f32)
    return x1, (mean1, rstd1, h1, qkv, o, lse)


def _self_attn_bwd(dx1, dx1b, x, n1w, wqkv, wproj, saved, B, L, H, meta, pairs=None):
    mean1, rstd1, h1, qkv, o, lse = saved
    D = x.shape[-1]
    dwproj = ops.linear_wgrad(dx1b, o)
    do = ops.linear_dgrad(dx1b, wproj)
    dqkv = torch.empty_like(qkv)
    ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, B, H, L, L, dqkv[:, :D], dqkv[:, D:2 * D],
                 dqkv[:, 2 * D:], meta=meta, pairs=pairs)
    dwqkv = ops.linear_wgrad(dqkv, h1)
    dh1 = ops.linear_dgrad(dqkv, wqkv)
    dn1w = _zeros(n1w.numel(), x.device)
    dx, dxb = ops.layernorm_bwd(dh1, x, n1w, mean1, rstd1, dx_in=dx1, d_weight=dn1w, want_bf16=True)
    return _with_bf16(dx, dxb), dn1w, dwqkv, dwproj


class _EncoderBlockFn(torch.autograd.Function):
    """x + attn(norm1 x) then + mlp(norm2 .) -- egom2p_utils.py:356-359."""

    @staticmethod
    def forward(ctx, x, n1w, qkv_w, proj_w, n2w, fc1_w, fc2_w, fc3_w, wb, geom):
        wqkv, wproj, w13, w2 = wb
        x1, sa = _self_attn_fwd(x, n1w, wqkv, wproj, geom.B, geom.N, geom.H, geom.m_enc, geom.eps, geom.p_enc)
        x2, sm = _mlp_fwd(x1, n2w, w13, w2, geom.eps)
        ctx.geom, ctx.wb, ctx.F = geom, wb, fc1_w.shape[0]
        ctx.op_epoch = geom.epoch_ref[0]
        ctx.save_for_backward(x, n1w, n2w, x1, *sa, *sm)
        return x2

    @staticmethod
    def backward(ctx, dx2):
        _GRAD_GEN[0] += 1
        _check_epoch(ctx.geom.epoch_ref, ctx.op_epoch)
        x, n1w, n2w, x1, *rest = ctx.saved_tensors
        sa, sm = rest[:6], rest[6:]
        wqkv, wproj, w13, w2 = ctx.wb
        g = ctx.geom
        dx2 = dx2.contiguous()
        dx1, dx1b, dn2w, dw13, dw2 = _mlp_bwd(dx2, x1, n2w, w13, w2, sm)
        dx, dn1w, dwqkv, dwproj = _self_attn_bwd(dx1, dx1b, x, n1w, wqkv, wproj, sa, g.B, g.N, g.H, g.m_enc, g.p_enc)
        F = ctx.F
        d1, d3 = _split_w13_grad(dw13, F)
        return dx, dn1w, dwqkv, dwproj, dn2w, d1, dw2[:, :F], d3, None, None


class _DecoderBlockFn(torch.autograd.Function):
    """self-attn -> cross-attn(query_norm y, context_norm ctx) -> mlp -- egom2p_utils.py:387-391."""

    @staticmethod
    def forward(ctx, y, context, n1w, qkv_w, sproj_w, qnw, cnw, q_w, kv_w, xproj_w, n2w, fc1_w, fc2_w, fc3_w, wb, geom):
        wqkv, wsproj, wq, wkv, wxproj, w13, w2 = wb
        g = geom
        D = y.shape[-1]
        y1, sa = _self_attn_fwd(y, n1w, wqkv, wsproj, g.B, g.M, g.H, g.m_dec, g.eps, g.p_dec)
        hq, _, meanq, rstdq = ops.layernorm_fwd(y1, qnw, g.eps)
        q = ops.linear_fwd(hq, wq)
        hc, _, meanc, rstdc = ops.layernorm_fwd(context, cnw, g.eps)
        kv = ops.linear_fwd(hc, wkv)
        o2, lse2 = ops.attn_fwd(q, kv[:, :D], kv[:, D:], g.B, g.H, g.M, g.N, meta=g.m_x, pairs=g.p_x)
        y2 = ops.linear_fwd(o2, wxproj, addend=y1, out_dtype=f32)
        y3, sm = _mlp_fwd(y2, n2w, w13, w2, g.eps)
        ctx.geom, ctx.wb, ctx.F = geom, wb, fc1_w.shape[0]
        ctx.op_epoch = geom.epoch_ref[0]
        ctx.ctx_key = context.data_ptr()
        g.ctx_users += 1
        ctx.save_for_backward(y, context, n1w, qnw, cnw, n2w, y1, y2, meanq, rstdq, hq, q, meanc, rstdc, hc, kv, o2, lse2,
                              *sa, *sm)
        return y3

    @staticmethod
    def backward(ctx, dy3):
        _GRAD_GEN[0] += 1
        _check_epoch(ctx.geom.epoch_ref, ctx.op_epoch)
        (y, context, n1w, qnw, cnw, n2w, y1, y2, meanq, rstdq, hq, q, meanc, rstdc, hc, kv, o2, lse2, *rest) = ctx.saved_tensors
        sa, sm = rest[:6], rest[6:]
        wqkv, wsproj, wq, wkv, wxproj, w13, w2 = ctx.wb
        g = ctx.geom
        D = y.shape[-1]
        dev = y.device
        dy3 = dy3.contiguous()
        dy2, dy2b, dn2w, dw13, dw2 = _mlp_bwd(dy3, y2, n2w, w13, w2, sm)
        # cross attention
        dwxproj = ops.linear_wgrad(dy2b, o2)
        do2 = ops.linear_dgrad(dy2b, wxproj)
        dq = torch.empty_like(q)
        dkv = torch.empty_like(kv)
        ops.attn_bwd(q, kv[:, :D], kv[:, D:], o2, do2, lse2, g.B, g.H, g.M, g.N, dq, dkv[:, :D], dkv[:, D:], meta=g.m_x, pairs=g.p_x)
        dwq = ops.linear_wgrad(dq, hq)
        dhq = ops.linear_dgrad(dq, wq)
        dwkv = ops.linear_wgrad(dkv, hc)
        dhc = ops.linear_dgrad(dkv, wkv)
        dqnw, dcnw = _zeros(D, dev), _zeros(D, dev)
        dy1, dy1b = ops.layernorm_bwd(dhq, y1, qnw, meanq, rstdq, dx_in=dy2, d_weight=dqnw, want_bf16=True)
        # All decoder blocks read the same `context`: instead of letting autograd add twelve (B*N, D) fp32 gradients pairwise,
        # each block's context_norm backward adds the running sum in its own pass (dx_in) and only the block that runs last
        # returns it. Falls back to one gradient per block whenever the bookkeeping does not match a plain single backward.
        acc = g.dctx_acc
        chained = g.ctx_users > 0 and (acc is None or (acc[0] == ctx.ctx_key and acc[1].shape == context.shape))
        last = chained and g.ctx_users == 1
        dctx, dctxb = ops.layernorm_bwd(dhc, context, cnw, meanc, rstdc, dx_in=acc[1] if (chained and acc is not None) else None,
                                        d_weight=dcnw, want_bf16=last)
        if chained:
            g.ctx_users -= 1
            if last:
                g.dctx_acc = None
                dctx = _with_bf16(dctx, dctxb)
            else:
                g.dctx_acc = (ctx.ctx_key, dctx)
                dctx = None
        dy, dn1w, dwqkv, dwsproj = _self_attn_bwd(dy1, dy1b, y, n1w, wqkv, wsproj, sa, g.B, g.M, g.H, g.m_dec, g.p_dec)
        F = ctx.F
        d1, d3 = _split_w13_grad(dw13, F)
        return (dy, dctx, dn1w, dwqkv, dwsproj, dqnw, dcnw, dwq, dwkv, dwxproj, dn2w, d1, dw2[:, :F], d3, None, None)


class _ContextFn(torch.autograd.Function):
    """context = decoder_proj_context(encoder_norm(x)) + encoder_emb -- egom2p_model.py:499,722."""

    @staticmethod
    def forward(ctx, x, enc_emb, norm_w, proj_w, proj_b, wb, eps, epoch_ref):
        h, _, mean, rstd = ops.layernorm_fwd(x, norm_w, eps)
        context = ops.linear_fwd(h, wb, bias=proj_b, addend=enc_emb, out_dtype=f32)
        ctx.wb = wb
        ctx.epoch_ref, ctx.op_epoch = epoch_ref, epoch_ref[0]
        ctx.save_for_backward(x, norm_w, mean, rstd, h)
        return context

    @staticmethod
    def backward(ctx, dctx):
        _GRAD_GEN[0] += 1
        _check_epoch(ctx.epoch_ref, ctx.op_epoch)
        x, norm_w, mean, rstd, h = ctx.saved_tensors
        dctx = dctx.contiguous()
        dcb = _bf16_of(dctx)
        dw = ops.linear_wgrad(dcb, h)
        db = ops.colsum(dctx, _zeros(dctx.shape[1], dctx.device))
        dh = ops.linear_dgrad(dcb, ctx.wb)
        dnw = _zeros(norm_w.numel(), x.device)
        dx, dxb = ops.layernorm_bwd(dh, x, norm_w, mean, rstd, d_weight=dnw, want_bf16=True)
        return _with_bf16(dx, dxb), dctx, dnw, dw, db, None, None, None


class _EmbedFn(torch.autograd.Function):
    """Fused token-embedding + pos/mod-embedding + masked gather for one side (north-star kernel 1).
    args: n token tables (encoder) or the mask token (decoder), then n modality embeddings."""

    @staticmethod
    def forward(ctx, plan, meta, is_decoder, *params):
        lens, vocabs, ids, pos, dim = meta
        n = len(lens)
        if is_decoder:
            mask_token, mods = params[0], params[1:]
            x0, emb = ops.embed_gather_fwd(plan, dim, lens, vocabs, None, None, pos, [m.reshape(-1) for m in mods],
                                           mask_token=mask_token.reshape(-1), want_emb=False)
        else:
            tables, mods = params[:n], params[n:]
            x0, emb = ops.embed_gather_fwd(plan, dim, lens, vocabs, ids, list(tables), pos, [m.reshape(-1) for m in mods])
        ctx.plan, ctx.meta, ctx.is_decoder = plan, meta, is_decoder
        ctx.shapes = [p.shape for p in params]
        ctx.save_for_backward(*[m for m in mods])
        R = plan.rows if plan.rows is not None else plan.B * plan.budget
        if is_decoder:
            return x0.reshape(R, dim)
        return x0.reshape(R, dim), emb.reshape(R, dim)

    @staticmethod
    def backward(ctx, dx0, demb=None):
        lens, vocabs, ids, pos, dim = ctx.meta
        n = len(lens)
        mods = ctx.saved_tensors
        dev = dx0.device
        dx0 = dx0.contiguous()
        d_mods = [_zeros(dim, dev) for _ in range(n)]
        if ctx.is_decoder:
            d_tok = _zeros(dim, dev)
            ops.embed_gather_bwd(ctx.plan, dim, lens, vocabs, None, pos, [m.reshape(-1) for m in mods], dx0, None, None,
                                 d_mods, d_tok)
            grads = [d_tok.reshape(ctx.shapes[0])] + [g.reshape(s) for g, s in zip(d_mods, ctx.shapes[1:])]
        else:
            d_tabs = [torch.zeros(s, dtype=f32, device=dev) for s in ctx.shapes[:n]]
            ops.embed_gather_bwd(ctx.plan, dim, lens, vocabs, ids, pos, [m.reshape(-1) for m in mods], dx0,
                                 demb.contiguous() if demb is not None else None, d_tabs, d_mods, None)
            grads = d_tabs + [g.reshape(s) for g, s in zip(d_mods, ctx.shapes[n:])]
        return (None, None, None, *grads)


class _HeadLossFn(torch.autograd.Function):
    """decoder_norm -> per-modality vocabulary head -> cross-entropy (egom2p_model.py:523,614-680) with the head GEMM
    fused into the softmax statistics: logits live in TMEM only. Returns the per-modality mean CE (0 for empty)."""
    CHUNK = 8192

    @staticmethod
    def forward(ctx, y, norm_w, rows, targets, wbs, eps, epoch_ref, *head_w):
        ctx.epoch_ref, ctx.op_epoch = epoch_ref, epoch_ref[0]
        yn, _, mean, rstd = ops.layernorm_fwd(y, norm_w, eps)
        losses, lses, yms = [], [], []
        for idx, tgt, wb in zip(rows, targets, wbs):
            if idx.numel() == 0:
                losses.append(torch.zeros((), dtype=f32, device=y.device))
                lses.append(None); yms.append(None)
                continue
            ym = ops.gather_rows_bf16(yn, idx)
            loss_sum, lse = ops.ce_forward(ym, wb, tgt)
            losses.append(loss_sum[0] / idx.numel())
            lses.append(lse); yms.append(ym)
        ctx.rows, ctx.targets, ctx.wbs, ctx.lses, ctx.yms = rows, targets, wbs, lses, yms
        ctx.save_for_backward(y, norm_w, mean, rstd)
        return tuple(losses)

    @staticmethod
    def backward(ctx, *dlosses):
        _GRAD_GEN[0] += 1
        _check_epoch(ctx.epoch_ref, ctx.op_epoch)
        y, norm_w, mean, rstd = ctx.saved_tensors
        dev = y.device
        D = y.shape[1]
        dyn = torch.zeros_like(y)  # rows of pads / modalities without targets get zero gradient
        dws = []
        for idx, tgt, wb, lse, ym, dl in zip(ctx.rows, ctx.targets, ctx.wbs, ctx.lses, ctx.yms, dlosses):
            V = wb.shape[0]
            if idx.numel() == 0 or dl is None:
                dws.append(torch.zeros(V, D, dtype=f32, device=dev))
                continue
            R = idx.numel()
            gscale = (dl.reshape(1).to(f32) / R).contiguous()
            dw = torch.empty(V, D, dtype=f32, device=dev)
            dym = torch.empty(R, D, dtype=f32, device=dev)
            chunk = min(V, _HeadLossFn.CHUNK)
            buf = torch.empty(R, chunk, dtype=bf16, device=dev)
            for v0 in range(0, V, chunk):
                vc = min(chunk, V - v0)
                dlog = buf[:, :vc]
                ops.ce_dlogits(ym, wb, tgt, lse, gscale, v0, vc, dlog)
                ops.gemm(dlog, wb[v0:v0 + vc], R, D, vc, b_mn=True, addend=dym if v0 else None, out_f32=dym)
                ops.gemm(dlog, ym, vc, D, R, a_mn=True, b_mn=True, out_f32=dw[v0:v0 + vc])
            ops.scatter_rows_f32(dym, idx, dyn)
            dws.append(dw)
        dnw = _zeros(D, dev)
        dy, dyb = ops.layernorm_bwd(dyn, y, norm_w, mean, rstd, d_weight=dnw, want_bf16=True)
        return (_with_bf16(dy, dyb), dnw, None, None, None, None, None, *dws)


# =============================================================================================== the module
class EgoM2P(nn.Module):
    """B200-native EgoM2P. Same constructor arguments as the reference (egom2p_model.py:84-108)."""

    def __init__(self,
                 encoder_embeddings: Dict[str, nn.Module],
                 decoder_embeddings: Dict[str, nn.Module],
                 modality_info: Dict[str, Any],
                 dim: int = 768,
                 encoder_depth: int = 12,
                 decoder_depth: int = 12,
                 num_heads: int = 12,
                 mlp_ratio: float = 4.0,
                 qkv_bias: bool = True,
                 proj_bias: bool = True,
                 mlp_bias: bool = True,
                 drop_path_rate_encoder: float = 0.0,
                 drop_path_rate_decoder: float = 0.0,
                 shared_drop_path: bool = False,
                 act_layer: nn.Module = nn.GELU,
                 norm_layer: Union[partial, nn.Module] = partial(LayerNorm, eps=1e-6, bias=False),
                 gated_mlp: bool = False,
                 qk_norm: bool = False,
                 decoder_causal_mask: bool = False,
                 decoder_sep_mask: bool = True,
                 num_register_tokens: int = 0,
                 use_act_checkpoint: bool = False,
                 share_modality_embeddings: bool = True):
        super().__init__()
        if qkv_bias or proj_bias or mlp_bias or not gated_mlp or qk_norm or act_layer is not nn.SiLU:
            raise NotImplementedError("egom2p_b200 implements the swiglu / no-bias family (egom2p_*_swiglu_nobias)")
        if drop_path_rate_encoder or drop_path_rate_decoder:
            raise NotImplementedError("drop_path > 0 is not built (the reference trains ego-b with 0)")
        if num_register_tokens:
            raise NotImplementedError("register tokens are not built (default 0 in the reference CLI)")
        if dim % num_heads or dim // num_heads != 64:
            raise NotImplementedError("attention kernels are specialised for head_dim 64 (tiny/small/base variants)")
        self.modality_info = modality_info
        self.dim, self.num_heads = dim, num_heads
        self.decoder_causal_mask, self.decoder_sep_mask = decoder_causal_mask, decoder_sep_mask
        self.init_std = 0.02
        self.use_act_checkpoint = use_act_checkpoint
        self.num_register_tokens = num_register_tokens
        eps_probe = norm_layer(4)
        self.eps = float(getattr(eps_probe, "eps", 1e-6))
        norm = partial(LayerNorm, eps=self.eps, bias=False)  # same names/buffers as the reference's LayerNorm

        self.encoder_modalities = set(encoder_embeddings.keys())
        for emb in encoder_embeddings.values():
            emb.init(dim_tokens=dim, init_std=self.init_std)
        self.encoder_embeddings = nn.ModuleDict(encoder_embeddings)
        self.decoder_modalities = set(decoder_embeddings.keys())
        for emb in decoder_embeddings.values():
            emb.init(dim_tokens=dim, init_std=self.init_std)
        self.decoder_embeddings = nn.ModuleDict(decoder_embeddings)
        for side in (self.encoder_embeddings, self.decoder_embeddings):
            for mod, emb in side.items():
                if isinstance(getattr(emb, "pos_emb", None), nn.Parameter):
                    # the fused embed backward emits gradients for the token tables, mask token and modality embeddings only
                    raise NotImplementedError(f"egom2p_b200: adapter '{mod}' has a learnable pos_emb (sincos_pos_emb=False); "
                                              "the fused path supports the fixed sin-cos tables ego-b uses")
        if share_modality_embeddings:
            self.share_modality_embeddings()

        self.encoder = nn.ModuleList([Block(dim, num_heads, mlp_ratio, norm) for _ in range(encoder_depth)])
        self.encoder_norm = norm(dim)
        self.decoder_proj_context = _Linear(dim, dim)
        self.decoder = nn.ModuleList([DecoderBlock(dim, num_heads, mlp_ratio, norm) for _ in range(decoder_depth)])
        self.decoder_norm = norm(dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, dim))
        nn.init.normal_(self.mask_token, std=self.init_std)
        self.register_tokens = None
        self.init_weights()
        self._wcache: Dict[Any, tuple] = {}
        self._wplan = None
        self._epoch = [0]   # bumped by every in-place refresh of the bf16 operands (see _check_epoch)
        self.static_target_rows: Optional[Dict[str, int]] = None   # {modality: valid target rows in the batch}, see forward
        self.fixed_decoder_order: Optional[List[str]] = None
        self._static_rows_dev = None
        self.pack_rows = True   # ragged batches: run every per-row kernel over the valid tokens only (see forward)

    # ------------------------------------------------------------------ construction helpers (reference :179-249)
    def share_modality_embeddings(self):
        for mod in self.encoder_modalities & self.decoder_modalities:
            self.decoder_embeddings[mod].mod_emb = self.encoder_embeddings[mod].mod_emb

    def init_weights(self):
        for name, m in self.named_modules():
            if "tokenizer" in name:
                continue
            if isinstance(m, nn.Linear):
                if "qkv" in name:
                    val = math.sqrt(6. / float(m.weight.shape[0] // 3 + m.weight.shape[1]))
                    nn.init.uniform_(m.weight, -val, val)
                elif "kv" in name:
                    val = math.sqrt(6. / float(m.weight.shape[0] // 2 + m.weight.shape[1]))
                    nn.init.uniform_(m.weight, -val, val)
                else:
                    nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.LayerNorm, LayerNorm)) or type(m).__name__ == "LayerNorm":
                nn.init.constant_(m.weight, 1.0)
                if getattr(m, "bias", None) is not None and isinstance(m.bias, nn.Parameter):
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=self.init_std)

    def get_num_layers_encoder(self):
        return len(self.encoder)

    def get_num_layers_decoder(self):
        return len(self.decoder)

    def get_num_layers(self):
        return self.get_num_layers_encoder() + self.get_num_layers_decoder()

    @torch.jit.ignore
    def no_weight_decay(self):
        no_wd = set()
        for side, embs in (("encoder_embeddings", self.encoder_embeddings), ("decoder_embeddings", self.decoder_embeddings)):
            for mod, emb in embs.items():
                if hasattr(emb, "no_weight_decay"):
                    no_wd |= {f"{side}.{mod}.{n}" for n in emb.no_weight_decay()}
        return no_wd

    # ------------------------------------------------------------------ bf16 operand cache (refreshed when a master changes)
    def _bf16(self, key, *params: torch.Tensor, pad32: bool = False) -> torch.Tensor:
        """bf16 GEMM operand of one weight (or of the fc1 | fc3 pair). Entries: (ptrs, versions, generation, operand, params)."""
        ptrs = tuple(p.data_ptr() for p in params)
        vers = tuple(p._version for p in params)
        hit = self._wcache.get(key)
        if hit is not None and hit[0] == ptrs:
            if hit[1] == vers and hit[2] == _GRAD_GEN[0]:
                return hit[3]
            self._refresh_operands()   # some master changed: one launch re-casts every registered operand
            hit = self._wcache[key]
            if hit[1] == vers and hit[2] == _GRAD_GEN[0]:
                return hit[3]
        dev = params[0].device
        if len(params) == 1:
            p = params[0]
            pad = (-p.shape[1]) % (32 if pad32 else 8)
            if pad == 0:
                wb = ops.cast_bf16(p.detach())
            else:  # inner dim padded with zero columns so TMA row pitches stay 16-byte multiples (e.g. hidden 682)
                wb = torch.zeros(p.shape[0], p.shape[1] + pad, dtype=bf16, device=dev)
                wb[:, :p.shape[1]].copy_(ops.cast_bf16(p.detach()))
        else:  # fc1 | fc3 -> one N = 2*hidden GEMM operand with rows interleaved in groups of 32 (hidden padded to 32)
            F, D = params[0].shape
            Fp = (F + 31) // 32 * 32
            wb = torch.zeros(2 * Fp, D, dtype=bf16, device=dev)
            v = wb.view(Fp // 32, 2, 32, D)
            for i, p in enumerate(params):
                tmp = ops.cast_bf16(p.detach())
                if Fp != F:
                    tmp = torch.cat([tmp, torch.zeros(Fp - F, D, dtype=bf16, device=dev)], 0)
                v[:, i].copy_(tmp.view(Fp // 32, 32, D))
        self._wcache[key] = (ptrs, vers, _GRAD_GEN[0], wb, params)
        self._wplan = None
        return wb

    def _refresh_operands(self):
        """Re-cast every cached operand whose masters still live at the same addresses, in one launch (ops.CastPlan)."""
        live = {k: e for k, e in self._wcache.items()
                if e[0] == tuple(p.data_ptr() for p in e[4]) and all(p.dim() == 2 and p.is_contiguous() for p in e[4])}
        if not live:
            return
        sig = tuple((k, e[0], e[3].data_ptr()) for k, e in live.items())
        if self._wplan is None or self._wplan[0] != sig:
            items = []
            for e in live.values():
                if len(e[4]) == 1:
                    items.append((e[4][0].detach(), e[3], 0, 0))
                else:
                    items += [(p.detach(), e[3], 32, i) for i, p in enumerate(e[4])]
            self._wplan = (sig, ops.CastPlan(items))
        self._wplan[1].run()
        self._epoch[0] += 1
        for k, e in live.items():
            self._wcache[k] = (e[0], tuple(p._version for p in e[4]), _GRAD_GEN[0], e[3], e[4])

    def invalidate_weight_cache(self):
        """Force a re-cast of every bf16 operand at the next forward. Needed only after an in-place weight update that
        neither bumps Tensor._version nor follows a backward of this module (see _GRAD_GEN)."""
        _GRAD_GEN[0] += 1

    def _enc_weights(self, i: int):
        b = self.encoder[i]
        return (self._bf16(("e", i, "qkv"), b.attn.qkv.weight), self._bf16(("e", i, "proj"), b.attn.proj.weight),
                self._bf16(("e", i, "w13"), b.mlp.fc1.weight, b.mlp.fc3.weight), self._bf16(("e", i, "w2"), b.mlp.fc2.weight, pad32=True))

    def _dec_weights(self, i: int):
        b = self.decoder[i]
        return (self._bf16(("d", i, "qkv"), b.self_attn.qkv.weight), self._bf16(("d", i, "sproj"), b.self_attn.proj.weight),
                self._bf16(("d", i, "q"), b.cross_attn.q.weight), self._bf16(("d", i, "kv"), b.cross_attn.kv.weight),
                self._bf16(("d", i, "xproj"), b.cross_attn.proj.weight),
                self._bf16(("d", i, "w13"), b.mlp.fc1.weight, b.mlp.fc3.weight), self._bf16(("d", i, "w2"), b.mlp.fc2.weight, pad32=True))

    # ------------------------------------------------------------------ sampler-facing pieces (reference :251-283,483-551)
    def cat_encoder_tensors(self, mod_dict):
        xs, es, ms, mm = [], [], [], []
        for mod, d in mod_dict.items():
            xs.append(d["x"]); es.append(d["emb"]); ms.append(d["input_mask"])
            mm.append(torch.full_like(d["input_mask"], self.modality_info[mod]["id"], dtype=torch.int16))
        return torch.cat(xs, dim=1), torch.cat(es, dim=1), torch.cat(ms, dim=1), torch.cat(mm, dim=1)

    @staticmethod
    def _prefix_ranges(mask: Optional[torch.Tensor], B: int, rows: int, keys: int, dev):
        """(B,1,keys) / (B,keys) key-padding mask or (B,rows,keys) mask (bool, True = masked) -> per-row [lo, hi) key
        ranges, int32 (B, rows). The masks this path meets are one contiguous run per row. The reduction
        runs on the mask's OWN shape (a key-padding mask costs O(B * keys), not O(B * rows * keys)) and nothing is read
        back: a non-contiguous row trips a device-side assertion (torch._assert_async) instead of a host sync."""
        if mask is None:
            return None, None
        m = mask.to(dev)
        if m.dim() == 2:
            m = m[:, None, :]
        if m.dtype != torch.bool:
            m = m != 0
        if m.shape[0] != B:
            m = m.expand(B, *m.shape[1:])
        r = m.shape[1]                     # 1 for a key-padding mask
        valid = (~m).to(torch.uint8)       # (B, r, keys), one byte per entry
        cnt = valid.sum(-1, dtype=torch.int32)
        first = valid.argmax(-1).to(torch.int32)                      # first valid key (0 when the row has none)
        last = keys - valid.flip(-1).argmax(-1).to(torch.int32)       # one past the last valid key
        torch._assert_async(((last - first == cnt) | (cnt == 0)).all(),
                            "egom2p_b200 attention supports one contiguous key range per query row")
        zero = torch.zeros_like(cnt)
        lo = torch.where(cnt > 0, first, zero)
        hi = torch.where(cnt > 0, last, zero)
        if r != rows:
            lo, hi = lo.expand(B, rows), hi.expand(B, rows)
        return lo.contiguous(), hi.contiguous()

    def forward_encoder(self, x: torch.Tensor, encoder_mask: torch.Tensor) -> torch.Tensor:
        B, N, D = x.shape
        if N == 0:
            return x
        g = self._geom(B, N, 0)
        g.enc_lo, g.enc_hi = self._prefix_ranges(encoder_mask, B, N, N, x.device)
        g.build_meta(x.device)
        h = x.reshape(B * N, D).float().contiguous()
        for i, blk in enumerate(self.encoder):
            h = _EncoderBlockFn.apply(h, blk.norm1.weight, blk.attn.qkv.weight, blk.attn.proj.weight, blk.norm2.weight,
                                      blk.mlp.fc1.weight, blk.mlp.fc2.weight, blk.mlp.fc3.weight, self._enc_weights(i), g)
        return self.encoder_norm(h).reshape(B, N, D)

    def forward_decoder(self, y, context, encoder_mask, decoder_attention_mask):
        B, M, D = y.shape
        N = context.shape[1]
        g = self._geom(B, N, M)
        g.x_lo, g.x_hi = self._prefix_ranges(encoder_mask, B, M, N, y.device) if N > 0 else (None, None)
        g.dec_lo, g.dec_hi = self._prefix_ranges(decoder_attention_mask, B, M, M, y.device)
        g.build_meta(y.device, encoder=False)  # no encoder self-attention on this path
        h = y.reshape(B * M, D).float().contiguous()
        c = context.reshape(B * N, D).float().contiguous() if N > 0 else torch.zeros(0, D, dtype=f32, device=y.device)
        if N == 0:
            # Unconditional branch of guided ROAR decoding (generate.py:793-802): no context token at all. Softmax over
            # zero keys times an empty V is exactly 0 and the cross-attention projection has no bias, so every block
            # reduces to self-attention + MLP. Inference only.
            if torch.is_grad_enabled() and (y.requires_grad or any(p.requires_grad for p in self.decoder.parameters())):
                raise NotImplementedError("decoder with an empty context is an inference-only path (wrap it in torch.no_grad())")
            for i, blk in enumerate(self.decoder):
                wqkv, wsproj, _, _, _, w13, w2 = self._dec_weights(i)
                h, _ = _self_attn_fwd(h, blk.norm1.weight, wqkv, wsproj, B, M, g.H, g.m_dec, g.eps)
                h, _ = _mlp_fwd(h, blk.norm2.weight, w13, w2, g.eps)
            return self.decoder_norm(h).reshape(B, M, D)
        for i, blk in enumerate(self.decoder):
            h = self._dec_block(i, blk, h, c, g)
        return self.decoder_norm(h).reshape(B, M, D)

    def _dec_block(self, i, blk, h, c, g):
        return _DecoderBlockFn.apply(h, c, blk.norm1.weight, blk.self_attn.qkv.weight, blk.self_attn.proj.weight,
                                     blk.query_norm.weight, blk.context_norm.weight, blk.cross_attn.q.weight,
                                     blk.cross_attn.kv.weight, blk.cross_attn.proj.weight, blk.norm2.weight,
                                     blk.mlp.fc1.weight, blk.mlp.fc2.weight, blk.mlp.fc3.weight, self._dec_weights(i), g)

    def forward_logits(self, y, decoder_mod_dict, decoder_mod_mask, return_all_logits: bool = False):
        out = {}
        for mod in decoder_mod_dict:
            idx = self.modality_info[mod]["id"]
            rows = y if return_all_logits else y[decoder_mod_mask == idx]
            out[mod] = self.decoder_embeddings[mod].forward_logits(rows)
        return out

    # ------------------------------------------------------------------ generation path (egom2p_b200/generate.py)
    @torch.no_grad()
    def encode_tokens(self, mod_dict, masks: Dict[str, torch.Tensor], budget: int):
        """Encoder half of one generation pass (reference: GenerationSampler.forward_mask_encoder_generation + forward_encoder
        + decoder_proj_context, generate.py:407-444,749-757) straight from token ids: index plan over `masks` (the
        modalities' input masks, possibly emptied for the unconditional branch), fused embed / gather of the `budget` kept
        slots, the encoder blocks and the context projection. Returns (context (B, budget, D) fp32, n_valid (B,) int32);
        no (B, L, D) embedding of the full modalities, no argsort, no host sync."""
        enc_mods = [m for m in mod_dict if m in self.encoder_embeddings]
        first = mod_dict[enc_mods[0]]["tensor"]
        B, dev, D = first.shape[0], first.device, self.dim
        ep = ops.index_plan([masks[m] for m in enc_mods], [self.modality_info[m]["id"] for m in enc_mods], budget)
        N = ep.budget
        lens, vocabs, ids, pos = self._tables("enc", enc_mods, mod_dict, B, dev)
        x, emb = ops.embed_gather_fwd(ep, D, lens, vocabs, ids, [self.encoder_embeddings[m].token_emb.weight.detach() for m in enc_mods],
                                      pos, [self.encoder_embeddings[m].mod_emb.detach().reshape(-1) for m in enc_mods])
        x, emb = x.reshape(B * N, D), emb.reshape(B * N, D)
        g = self._geom(B, N, 0)
        g.enc_lo = torch.zeros(B, N, dtype=torch.int32, device=dev)
        g.enc_hi = ep.n_valid[:, None].expand(B, N).contiguous()
        g.build_meta(dev)
        for i, blk in enumerate(self.encoder):
            wqkv, wproj, w13, w2 = self._enc_weights(i)
            x, _ = _self_attn_fwd(x, blk.norm1.weight.detach(), wqkv, wproj, B, N, g.H, g.m_enc, g.eps)
            x, _ = _mlp_fwd(x, blk.norm2.weight.detach(), w13, w2, g.eps)
        h, _, _, _ = ops.layernorm_fwd(x, self.encoder_norm.weight.detach(), self.eps, save_stats=False)
        context = ops.linear_fwd(h, self._bf16("ctx", self.decoder_proj_context.weight), bias=self.decoder_proj_context.bias.detach(),
                                 addend=emb, out_dtype=f32)
        return context.reshape(B, N, D), ep.n_valid

    @torch.no_grad()
    def decode_tokens(self, y0: torch.Tensor, context: Optional[torch.Tensor], n_ctx: Optional[torch.Tensor]):
        """Decoder half of one generation pass (generate.py:758-763 with decoder_attention_mask=None): y0 (B, k, D) fp32 decoder
        inputs, context (B, N, D) fp32 with the first n_ctx[b] rows of sample b valid (a sample with n_ctx = 0 has NO context:
        its cross-attention contributes exactly 0, SURVEY A5 (i) -- that is what lets the conditional and the unconditional
        branch of guided decoding share one batch). Returns decoder_norm(y) as (B * k, D) fp32."""
        B, M, D = y0.shape
        dev = y0.device
        N = 0 if context is None else context.shape[1]
        g = self._geom(B, N, M)
        if N > 0:
            g.x_lo = torch.zeros(B, M, dtype=torch.int32, device=dev)
            g.x_hi = n_ctx.to(torch.int32)[:, None].expand(B, M).contiguous()
            g.m_x = ops.attn_ranges(B, M, N, g.x_lo, g.x_hi, device=dev, empty_zero=True)
            c = context.reshape(B * N, D)
        g.m_dec = ops.attn_ranges(B, M, M, device=dev)
        h = y0.reshape(B * M, D).float().contiguous()
        for i, blk in enumerate(self.decoder):
            wqkv, wsproj, wq, wkv, wxproj, w13, w2 = self._dec_weights(i)
            h, _ = _self_attn_fwd(h, blk.norm1.weight.detach(), wqkv, wsproj, B, M, g.H, g.m_dec, g.eps)
            if N > 0:
                hq, _, _, _ = ops.layernorm_fwd(h, blk.query_norm.weight.detach(), g.eps, save_stats=False)
                q = ops.linear_fwd(hq, wq)
                hc, _, _, _ = ops.layernorm_fwd(c, blk.context_norm.weight.detach(), g.eps, save_stats=False)
                kv = ops.linear_fwd(hc, wkv)
                o2, _ = ops.attn_fwd(q, kv[:, :D], kv[:, D:], B, g.H, M, N, meta=g.m_x, want_lse=False)
                h = ops.linear_fwd(o2, wxproj, addend=h, out_dtype=f32)
            h, _ = _mlp_fwd(h, blk.norm2.weight.detach(), w13, w2, g.eps)
        _, yn, _, _ = ops.layernorm_fwd(h, self.decoder_norm.weight.detach(), self.eps, out_bf16=False, out_f32=True, save_stats=False)
        return yn

    def head_operand(self, mod: str) -> torch.Tensor:
        """Cached bf16 operand of a modality's vocabulary head (V, D)."""
        return self._bf16(("head", mod), self.decoder_embeddings[mod].to_logits.weight)

    @staticmethod
    def _pack_plan(pl: "ops.Plan", n_valid: List[int], budget: int):
        """Packed view of an index plan: the valid slots only (a prefix of every sample's slots), sample after sample.
        Returns (plan over the packed rows, row offset of every sample, valid rows of every sample) -- device tensors built
        from the host counts, no further synchronisation."""
        dev = pl.keep_mod.device
        B = pl.B
        lens = torch.tensor(n_valid, dtype=torch.int32, device=dev)
        off = torch.cumsum(lens, 0, dtype=torch.int32) - lens
        R = int(sum(n_valid))
        ar = torch.arange(R, dtype=torch.int32, device=dev)
        b_of = torch.searchsorted((off + lens).contiguous(), ar, right=True).to(torch.int32)
        slot = (b_of.long() * budget + (ar - off[b_of.long()]).long())          # flat (b, s) index of every packed row
        pp = ops.Plan()
        pp.B, pp.budget, pp.rows, pp.row_batch = B, budget, R, b_of.contiguous()
        pp.keep_mod = pl.keep_mod.reshape(-1)[slot].contiguous()
        pp.keep_pos = pl.keep_pos.reshape(-1)[slot].contiguous()
        pp.pad = torch.zeros(R, dtype=torch.bool, device=dev)
        pp.mod_mask = pl.mod_mask.reshape(-1)[slot]
        pp.n_valid = pl.n_valid
        pp.keep_idx = None
        pp.target_ids = pp.key_lo = pp.key_hi = None
        if pl.target_ids is not None:
            pp.target_ids = pl.target_ids          # indexed by SLOT (the head row lists are slot lists before remapping)
            pp.key_lo = pl.key_lo.reshape(-1)[slot]
            pp.key_hi = pl.key_hi.reshape(-1)[slot]
        pp.slot_index = slot
        pp.slot_to_row = torch.full((B * budget,), -1, dtype=torch.int64, device=dev)
        pp.slot_to_row[slot] = torch.arange(R, dtype=torch.int64, device=dev)
        return pp, off, lens

    def _geom(self, B, N, M) -> _Geom:
        g = _Geom()
        g.B, g.N, g.M, g.H, g.D, g.eps = B, N, M, self.num_heads, self.dim, self.eps
        g.enc_lo = g.enc_hi = g.dec_lo = g.dec_hi = g.x_lo = g.x_hi = None
        g.m_enc = g.m_dec = g.m_x = None
        g.ctx_users, g.dctx_acc = 0, None   # decoder blocks of this forward that read `context` / their running gradient
        g.epoch_ref = self._epoch
        return g

    # ------------------------------------------------------------------ the training step (reference :683-734)
    def _tables(self, side: str, mods: List[str], mod_dict, B, dev):
        embs = self.encoder_embeddings if side == "enc" else self.decoder_embeddings
        lens, vocabs, ids, pos = [], [], [], []
        for m in mods:
            e = embs[m]
            t = mod_dict[m]["tensor"].reshape(B, -1)
            if t.dtype != torch.int64:
                t = t.to(torch.int64)
            ids.append(t.contiguous())
            lens.append(t.shape[1])
            vocabs.append(int(e.vocab_size))
            pos.append(e.pos_emb.detach().reshape(-1, self.dim))
        return lens, vocabs, ids, pos

    def forward(self, mod_dict: Dict[str, Dict[str, torch.Tensor]], num_encoder_tokens: int, num_decoder_tokens: int,
                loss_type: str = "mod", return_logits: bool = False):
        if loss_type not in ("mod", "modality", "weighted_mod", "token"):
            raise ValueError("Invalid loss type")
        enc_mods = [m for m in mod_dict if m in self.encoder_embeddings]
        dec_mods = [m for m in mod_dict if m in self.decoder_embeddings]
        first = mod_dict[enc_mods[0]]["tensor"]
        if not first.is_cuda:
            raise RuntimeError("egom2p_b200: mod_dict tensors must live on a CUDA device (no CPU path exists)")
        B, dev, D = first.shape[0], first.device, self.dim
        ids_of = lambda m: self.modality_info[m]["id"]

        # ---- index plans (no embedding rows touched; masks -> slots, pads, key ranges)
        ep = ops.index_plan([mod_dict[m]["input_mask"] for m in enc_mods], [ids_of(m) for m in enc_mods], num_encoder_tokens)
        # decoder modality order is shuffled exactly like the reference (random.sample over the dict items, :312)
        dec_order = [m for m, _ in random.sample([(m, None) for m in dec_mods], len(dec_mods))]
        if self.fixed_decoder_order is not None:   # CUDA-graph replay: the order drawn at capture time (the loss is order-invariant, A7)
            dec_order = [m for m in self.fixed_decoder_order if m in dec_mods]
        for m in dec_order:
            if self.modality_info[m]["type"] in ("seq", "seq_emb", "seq_token"):
                raise NotImplementedError("sequence (teacher-forced) decoder modalities are not part of the mod4 path")
        dlens, dvocabs, dids, dpos = self._tables("dec", dec_order, mod_dict, B, dev)
        dp = ops.index_plan([mod_dict[m]["target_mask"] for m in dec_order], [ids_of(m) for m in dec_order], num_decoder_tokens,
                            decoder=True, attn_cnt=[mod_dict[m]["decoder_attention_mask"].to(torch.int32) for m in dec_order],
                            ids=dids, causal=self.decoder_causal_mask, sep=self.decoder_sep_mask)
        N, M = ep.budget, dp.budget
        # ---- row lists of the heads, and ONE device -> host copy per forward: valid tokens per sample on both sides and the
        # rows per head (the reference's boolean row-selects sync once per modality, egom2p_model.py:633). With
        # `static_target_rows` set (CUDA-graph capture, egom2p_b200/graphed.py) nothing is copied.
        row_tab, counts = ops.plan_rows(dp.mod_mask, [ids_of(m) for m in dec_mods])
        pack = None
        if self.static_target_rows is not None:
            n_rows = [int(self.static_target_rows[m]) for m in dec_mods]
            key = (tuple(n_rows), dev)
            if self._static_rows_dev is None or self._static_rows_dev[0] != key:   # built outside graph capture (warm-up)
                self._static_rows_dev = (key, torch.tensor(n_rows, dtype=torch.int32, device=dev))
            torch._assert_async((counts == self._static_rows_dev[1]).all(),
                                "egom2p_b200: static_target_rows does not match this batch")
        else:
            host = torch.cat([ep.n_valid, dp.n_valid, counts]).tolist()
            ne, nd, n_rows = host[:B], host[B:2 * B], host[2 * B:]
            # Ragged batch: PACK the rows (valid slots only, sample after sample) so that LayerNorm, every GEMM and the
            # attention queries run over the valid tokens instead of the full budgets (the reference's masks leave ~48 % of
            # the 2048 + 2048 slots valid on average). A sample without any encoder token keeps the
            # slot layout: its cross-attention must average over its own pad keys.
            if self.pack_rows and min(ne) > 0 and (sum(ne) < B * N or sum(nd) < B * M) and sum(nd) > 0:
                pack = (ne, nd)
        if pack is not None:
            ep_p, e_off, e_len = self._pack_plan(ep, pack[0], N)
            dp_p, d_off, d_len = self._pack_plan(dp, pack[1], M)
            Re, Rd = ep_p.rows, dp_p.rows
            g = self._geom(1, Re, Rd)        # one virtual sample: block-diagonal key ranges keep the samples apart
            eb, db = ep_p.row_batch.long(), dp_p.row_batch.long()
            g.enc_lo = e_off[eb].reshape(1, Re).contiguous()
            g.enc_hi = (e_off[eb] + e_len[eb]).reshape(1, Re).contiguous()
            g.x_lo = e_off[db].reshape(1, Rd).contiguous()
            g.x_hi = (e_off[db] + e_len[db]).reshape(1, Rd).contiguous()
            g.dec_lo = (dp_p.key_lo + d_off[db]).reshape(1, Rd).contiguous()
            g.dec_hi = (dp_p.key_hi + d_off[db]).reshape(1, Rd).contiguous()
            ep, dp = ep_p, dp_p
        else:
            g = self._geom(B, N, M)
            nv = ep.n_valid
            g.enc_lo = torch.zeros(B, N, dtype=torch.int32, device=dev)
            g.enc_hi = nv[:, None].expand(B, N).contiguous()
            g.x_lo = torch.zeros(B, M, dtype=torch.int32, device=dev)
            g.x_hi = nv[:, None].expand(B, M).contiguous()
            g.dec_lo, g.dec_hi = dp.key_lo, dp.key_hi
        g.build_meta(dev)

        # ---- fused embed / gather
        elens, evocabs, eids, epos = self._tables("enc", enc_mods, mod_dict, B, dev)
        x, enc_emb = _EmbedFn.apply(ep, (elens, evocabs, eids, epos, D), False,
                                    *[self.encoder_embeddings[m].token_emb.weight for m in enc_mods],
                                    *[self.encoder_embeddings[m].mod_emb for m in enc_mods])
        y = _EmbedFn.apply(dp, (dlens, dvocabs, None, dpos, D), True, self.mask_token,
                           *[self.decoder_embeddings[m].mod_emb for m in dec_order])

        # ---- encoder
        for i, blk in enumerate(self.encoder):
            x = _EncoderBlockFn.apply(x, blk.norm1.weight, blk.attn.qkv.weight, blk.attn.proj.weight, blk.norm2.weight,
                                      blk.mlp.fc1.weight, blk.mlp.fc2.weight, blk.mlp.fc3.weight, self._enc_weights(i), g)
        context = _ContextFn.apply(x, enc_emb, self.encoder_norm.weight, self.decoder_proj_context.weight,
                                   self.decoder_proj_context.bias, self._bf16("ctx", self.decoder_proj_context.weight), self.eps,
                                   self._epoch)
        # ---- decoder
        for i, blk in enumerate(self.decoder):
            y = self._dec_block(i, blk, y, context, g)

        if return_logits:
            yn = self.decoder_norm(y)
            if pack is not None:   # back to the (B, M) slot layout the reference returns (pad slots: zeros)
                full = torch.zeros(B * M, D, dtype=yn.dtype, device=dev)
                full[dp.slot_index] = yn
                yn = full
            yn = yn.reshape(B, M, D)
            return {m: self.decoder_embeddings[m].forward_logits(yn) for m in dec_mods}

        # ---- heads + loss (row lists from plan_rows above; in the packed layout the slot indices are remapped)
        tgt_flat = dp.target_ids.reshape(-1)          # indexed by slot in both layouts
        rows, targets, wbs, heads = [], [], [], []
        for i, m in enumerate(dec_mods):
            idx = row_tab[i, :n_rows[i]]              # slots (b * M + s) of this modality's valid targets, ascending
            targets.append(tgt_flat.index_select(0, idx))
            rows.append(dp.slot_to_row.index_select(0, idx) if pack is not None else idx)
            w = self.decoder_embeddings[m].to_logits.weight
            heads.append(w)
            wbs.append(self._bf16(("head", m), w))
        per_mod = _HeadLossFn.apply(y, self.decoder_norm.weight, rows, targets, wbs, self.eps, self._epoch, *heads)
        mod_loss = {}
        for m, l in zip(dec_mods, per_mod):
            if loss_type == "weighted_mod" and rows[dec_mods.index(m)].numel():
                l = l / math.log(self.modality_info[m]["vocab_size"]) * 5.545177444479562
            mod_loss[m] = l
        if loss_type == "token":
            counts = {m: rows[i].numel() * int(self.decoder_embeddings[m].vocab_size) for i, m in enumerate(dec_mods)}
            loss = sum(mod_loss[m] * counts[m] for m in dec_mods) / sum(counts.values())
        else:
            loss = sum(mod_loss.values()) / len(mod_loss)
        return loss, mod_loss

    # ------------------------------------------------------------------ freeze helpers (reference :737-819)
    def _set_grad(self, modules, flag):
        for mod in modules:
            for p in mod.parameters():
                p.requires_grad = flag

    def freeze_encoder(self, freeze_embeddings=True):
        self._set_grad([self.encoder, self.encoder_norm] + ([self.encoder_embeddings] if freeze_embeddings else []), False)

    def freeze_encoder_except_specific_embeddings(self, frozen_embedding_domain):
        doms = frozen_embedding_domain.split("-")
        self._set_grad([self.encoder, self.encoder_norm], False)
        for name, p in self.encoder_embeddings.named_parameters():
            if name.split(".")[0] in doms:
                p.requires_grad = False

    def unfreeze_encoder(self, unfreeze_embeddings=True):
        self._set_grad([self.encoder, self.encoder_norm] + ([self.encoder_embeddings] if unfreeze_embeddings else []), True)

    def freeze_decoder(self, freeze_embeddings=True):
        self._set_grad([self.decoder, self.decoder_norm] + ([self.decoder_embeddings] if freeze_embeddings else []), False)

    def freeze_decoder_except_specific_embeddings(self, frozen_embedding_domain):
        doms = frozen_embedding_domain.split("-")
        self._set_grad([self.decoder, self.decoder_norm], False)
        for name, p in self.decoder_embeddings.named_parameters():
            if name.split(".")[0] in doms:
                p.requires_grad = False

    def unfreeze_decoder(self, unfreeze_embeddings=True):
        self._set_grad([self.decoder, self.decoder_norm] + ([self.decoder_embeddings] if unfreeze_embeddings else []), True)

    def freeze_shared_params(self):
        self.freeze_encoder(freeze_embeddings=False)
        self.freeze_decoder(freeze_embeddings=False)

    def freeze_params_except_specific_embeddings(self, frozen_embedding_domain):
        self.freeze_encoder_except_specific_embeddings(frozen_embedding_domain)
        self.freeze_decoder_except_specific_embeddings(frozen_embedding_domain)

    def unfreeze_shared_params(self):
        self.unfreeze_encoder(unfreeze_embeddings=False)
        self.unfreeze_decoder(unfreeze_embeddings=False)

    def unfreeze_all(self):
        self.unfreeze_encoder(unfreeze_embeddings=True)
        self.unfreeze_decoder(unfreeze_embeddings=True)
