"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (egom2p_b200/).

CPU fp32 restatement of the reference EgoM2P masked multimodal training step, written from the
reference's behaviour (files cited per function, paths relative to /root/reference). It is the
checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it.

Parity status: PINNED. `oracle/gen_golden.py` imports the unmodified reference in the authoring
container, runs it on seeded inputs and stores its outputs under tests/golden/; tests/test_oracle.py
checks this restatement against those fixtures (index plan bit-exact, loss/logits/grads to fp32
round-off). The reference itself ships no tests or golden vectors.

Scope: the swiglu/no-bias family used by ego-b (`egom2p_*_swiglu_nobias`), modalities of type
img/cam/gaze (token ids in, masked-token decoder inputs).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# positional tables  (egom2p/models/egom2p_utils.py:32-44, 63-86)
# ----------------------------------------------------------------------------------------------
def sincos_1d(n: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    assert dim % 2 == 0
    pos = torch.arange(n, dtype=torch.float32)
    half = dim // 2
    omega = 1.0 / (temperature ** (torch.arange(half, dtype=torch.float32) / half))
    out = pos[:, None] * omega[None, :]
    return torch.cat([out.sin(), out.cos()], dim=1)  # (n, dim)


def sincos_3d(t: int, h: int, w: int, dim: int, temperature: float = 10000.0) -> torch.Tensor:
    assert dim % 6 == 0
    ch = dim // 6 * 2
    inv = 1.0 / (temperature ** (torch.arange(0, ch, 2).float() / ch))

    def axis(n):
        a = torch.arange(n, dtype=torch.float32)[:, None] * inv[None, :]
        return torch.stack((a.sin(), a.cos()), dim=-1).flatten(-2, -1)  # (n, ch) interleaved sin/cos

    et, eh, ew = axis(t), axis(h), axis(w)
    emb = torch.zeros(t, h, w, 3 * ch)
    emb[..., :ch] = et[:, None, None, :]
    emb[..., ch:2 * ch] = eh[None, :, None, :]
    emb[..., 2 * ch:] = ew[None, None, :, :]
    return emb.reshape(t * h * w, dim)


# ----------------------------------------------------------------------------------------------
# index plan  (egom2p/models/egom2p_model.py:251-283, 344-396, 398-481)
# ----------------------------------------------------------------------------------------------
def _keep_ids(mask_all: np.ndarray, budget: int) -> np.ndarray:
    """ids_keep = argsort(mask + arange*1e-6)[:, :budget], fp32 keys (egom2p_model.py:370-373)."""
    L = mask_all.shape[1]
    keys = mask_all.astype(np.float32) + (np.arange(L, dtype=np.int64) * 1e-6).astype(np.float32)[None, :]
    return np.argsort(keys, axis=1, kind="stable")[:, :budget]


def plan_encoder(input_masks: Dict[str, np.ndarray], mod_ids: Dict[str, int], budget: int):
    """forward_mask_encoder index part. `input_masks` in mod_dict order; True = not an input."""
    mask_all = np.concatenate([m.reshape(m.shape[0], -1) for m in input_masks.values()], axis=1)
    mod_all = np.concatenate([np.full(m.reshape(m.shape[0], -1).shape, mod_ids[k], dtype=np.int16)
                              for k, m in input_masks.items()], axis=1)
    ids_keep = _keep_ids(mask_all, budget)
    mask = np.take_along_axis(mask_all, ids_keep, axis=1)
    mod = np.take_along_axis(mod_all, ids_keep, axis=1).copy()
    mod[mask] = -1
    return {"ids_keep": ids_keep, "mask": mask, "mod_mask": mod}


def plan_decoder(target_masks: Dict[str, np.ndarray], attn_counts: Dict[str, np.ndarray],
                 ids: Dict[str, np.ndarray], mod_ids: Dict[str, int], order: List[str], budget: int,
                 causal: bool = False, sep: bool = True):
    """forward_mask_decoder index part; `order` is the (shuffled) modality order the reference draws
    with random.sample (egom2p_model.py:312)."""
    B = next(iter(target_masks.values())).shape[0]
    mask_all = np.concatenate([target_masks[k].reshape(B, -1) for k in order], axis=1)
    cnt_all = np.concatenate([attn_counts[k].reshape(B, -1) for k in order], axis=1).astype(np.int32)
    ids_all = np.concatenate([ids[k].reshape(B, -1) for k in order], axis=1).astype(np.int64)
    mod_all = np.concatenate([np.full((B, target_masks[k].reshape(B, -1).shape[1]), mod_ids[k], dtype=np.int16)
                              for k in order], axis=1)
    ids_keep = _keep_ids(mask_all, budget)
    mask = np.take_along_axis(mask_all, ids_keep, axis=1)
    tgt = np.take_along_axis(ids_all, ids_keep, axis=1).copy()
    cnt = np.take_along_axis(cnt_all, ids_keep, axis=1)
    mod = np.take_along_axis(mod_all, ids_keep, axis=1).copy()
    tgt[mask] = 0
    M = ids_keep.shape[1]
    # adapt_decoder_attention_mask (egom2p_model.py:446-481), evaluated BEFORE mod ids become -1
    if causal:
        attn = np.broadcast_to(np.triu(np.ones((M, M), dtype=bool), 1), (B, M, M)).copy()
    else:
        cum = np.cumsum(cnt, axis=1)
        attn = np.arange(M)[None, None, :] >= cum[:, :, None]
    if sep:
        attn = attn | (mod[:, None, :] != mod[:, :, None])
    mod[mask] = -1
    return {"ids_keep": ids_keep, "mask": mask, "mod_mask": mod, "target_ids": tgt, "attn_mask": attn}


# ----------------------------------------------------------------------------------------------
# blocks  (egom2p/models/egom2p_utils.py:118-133, 154-244, 335-391)
# ----------------------------------------------------------------------------------------------
def _ln(x, w, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), w, None, eps)


def _attend(q, k, v, mask, heads):
    """softmax((q k^T) d^-1/2 masked_fill(mask, -finfo.max)) v  (egom2p_utils.py:190-202, 230-241).
    q (B,Nq,D), k/v (B,Nk,D), mask broadcastable to (B,Nq,Nk) or None; True = masked."""
    B, Nq, D = q.shape
    Nk = k.shape[1]
    d = D // heads
    q = q.reshape(B, Nq, heads, d).permute(0, 2, 1, 3)
    k = k.reshape(B, Nk, heads, d).permute(0, 2, 1, 3)
    v = v.reshape(B, Nk, heads, d).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-2, -1)) * (d ** -0.5)
    if mask is not None:
        s = s.masked_fill(mask[:, None], -torch.finfo(s.dtype).max)
    p = s.softmax(dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, Nq, D)


def _mlp(x, sd, pre):
    return F.linear(F.silu(F.linear(x, sd[pre + "fc1.weight"])) * F.linear(x, sd[pre + "fc3.weight"]),
                    sd[pre + "fc2.weight"])


def encoder_block(x, sd, pre, mask, heads):
    h = _ln(x, sd[pre + "norm1.weight"])
    qkv = F.linear(h, sd[pre + "attn.qkv.weight"])
    D = x.shape[-1]
    a = _attend(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], mask, heads)
    x = x + F.linear(a, sd[pre + "attn.proj.weight"])
    return x + _mlp(_ln(x, sd[pre + "norm2.weight"]), sd, pre + "mlp.")


def decoder_block(y, ctx, sd, pre, sa_mask, xa_mask, heads):
    D = y.shape[-1]
    h = _ln(y, sd[pre + "norm1.weight"])
    qkv = F.linear(h, sd[pre + "self_attn.qkv.weight"])
    a = _attend(qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:], sa_mask, heads)
    y = y + F.linear(a, sd[pre + "self_attn.proj.weight"])
    q = F.linear(_ln(y, sd[pre + "query_norm.weight"]), sd[pre + "cross_attn.q.weight"])
    kv = F.linear(_ln(ctx, sd[pre + "context_norm.weight"]), sd[pre + "cross_attn.kv.weight"])
    a = _attend(q, kv[..., :D], kv[..., D:], xa_mask, heads)
    y = y + F.linear(a, sd[pre + "cross_attn.proj.weight"])
    return y + _mlp(_ln(y, sd[pre + "norm2.weight"]), sd, pre + "mlp.")


def forward_encoder(x, enc_mask, sd, depth, heads):
    """egom2p_model.py:483-501. enc_mask (B,1,N) bool, True = pad key."""
    for i in range(depth):
        x = encoder_block(x, sd, f"encoder.{i}.", enc_mask, heads)
    return _ln(x, sd["encoder_norm.weight"])


def forward_decoder(y, ctx, enc_mask, dec_attn_mask, sd, depth, heads):
    """egom2p_model.py:503-525."""
    for i in range(depth):
        y = decoder_block(y, ctx, sd, f"decoder.{i}.", dec_attn_mask, enc_mask, heads)
    return _ln(y, sd["decoder_norm.weight"])


# ----------------------------------------------------------------------------------------------
# whole step  (egom2p/models/egom2p_model.py:683-734; adapters encoder_embeddings.py:181-301,
# decoder_embeddings.py:337-500)
# ----------------------------------------------------------------------------------------------
def forward(sd: Dict[str, torch.Tensor], cfg: dict, mod_dict: Dict[str, Dict[str, torch.Tensor]],
            num_encoder_tokens: int, num_decoder_tokens: int, dec_order: Optional[List[str]] = None,
            loss_type: str = "mod", return_logits: bool = False, keep: bool = False):
    """cfg: {dim, heads, enc_depth, dec_depth, mods: {name: {id, vocab}}, causal?, sep?}.
    `sd` uses the reference state_dict keys. Returns a dict with loss, mod_loss and (keep=True) the
    intermediates the parity tests compare (plans, gathered embeddings, encoder/decoder outputs)."""
    heads, D = cfg["heads"], cfg["dim"]
    mods = [m for m in mod_dict if m in cfg["mods"]]
    mod_ids = {m: cfg["mods"][m]["id"] for m in mods}
    B = mod_dict[mods[0]]["tensor"].shape[0]
    dec_order = dec_order or mods

    # --- encoder side: x = token_emb[ids], emb = pos_emb + mod_emb; cat, keep, zero pads
    ep = plan_encoder({m: mod_dict[m]["input_mask"].numpy() for m in mods}, mod_ids, num_encoder_tokens)
    xs, es = [], []
    for m in mods:
        ids = mod_dict[m]["tensor"].reshape(B, -1).long()
        pre = f"encoder_embeddings.{m}."
        xs.append(sd[pre + "token_emb.weight"][ids])
        es.append((sd[pre + "pos_emb"] + sd[pre + "mod_emb"]).expand(B, -1, -1))
    x_all, e_all = torch.cat(xs, 1), torch.cat(es, 1)
    keep_e = torch.from_numpy(ep["ids_keep"])[..., None].expand(-1, -1, D)
    enc_tok = torch.gather(x_all, 1, keep_e)
    enc_emb = torch.gather(e_all, 1, keep_e)
    emask = torch.from_numpy(ep["mask"])
    enc_tok = enc_tok.masked_fill(emask[..., None], 0.0)
    enc_emb = enc_emb.masked_fill(emask[..., None], 0.0)

    # --- decoder side: tokens := mask_token, emb = pos_emb + mod_emb, in dec_order
    dp = plan_decoder({m: mod_dict[m]["target_mask"].numpy() for m in mods},
                      {m: mod_dict[m]["decoder_attention_mask"].numpy() for m in mods},
                      {m: mod_dict[m]["tensor"].reshape(B, -1).numpy() for m in mods},
                      mod_ids, dec_order, num_decoder_tokens,
                      causal=cfg.get("causal", False), sep=cfg.get("sep", True))
    es = []
    for m in dec_order:
        pre = f"decoder_embeddings.{m}."
        es.append((sd[pre + "pos_emb"] + sd[pre + "mod_emb"]).expand(B, -1, -1))
    e_all = torch.cat(es, 1)
    keep_d = torch.from_numpy(dp["ids_keep"])[..., None].expand(-1, -1, D)
    dec_emb = torch.gather(e_all, 1, keep_d)
    dmask = torch.from_numpy(dp["mask"])
    dec_tok = (torch.zeros_like(dec_emb) + sd["mask_token"]).masked_fill(dmask[..., None], 0.0)
    dec_emb = dec_emb.masked_fill(dmask[..., None], 0.0)

    x = enc_tok + enc_emb
    enc_mask3 = emask[:, None, :]
    x = forward_encoder(x, enc_mask3, sd, cfg["enc_depth"], heads)
    ctx = F.linear(x, sd["decoder_proj_context.weight"], sd["decoder_proj_context.bias"]) + enc_emb
    y0 = dec_tok + dec_emb
    y = forward_decoder(y0, ctx, enc_mask3, torch.from_numpy(dp["attn_mask"]), sd, cfg["dec_depth"], heads)

    out = {}
    if keep:
        out.update(enc_plan=ep, dec_plan=dp, enc_x0=enc_tok + enc_emb, enc_emb=enc_emb, dec_y0=y0,
                   enc_out=x, context=ctx, dec_out=y)
    dec_mod = torch.from_numpy(dp["mod_mask"].astype(np.int64))
    tgt = torch.from_numpy(dp["target_ids"])
    if return_logits:
        out["logits"] = {m: F.linear(y, sd[f"decoder_embeddings.{m}.to_logits.weight"]) for m in mods}
        return out
    # forward_mod_loss / forward_token_loss / forward_weighted_mod_loss (egom2p_model.py:581-680)
    mod_loss, mod_count = {}, {}
    for m in mods:
        sel = dec_mod == mod_ids[m]
        logits = F.linear(y[sel], sd[f"decoder_embeddings.{m}.to_logits.weight"])
        if logits.numel() == 0:
            mod_loss[m] = logits.sum()
            mod_count[m] = 0
        else:
            ce = F.cross_entropy(logits, tgt[sel], reduction="mean")
            if loss_type == "weighted_mod":
                ce = ce / math.log(cfg["mods"][m]["vocab"]) * 5.545177444479562
            mod_loss[m] = ce
            mod_count[m] = logits.numel()
    if loss_type in ("mod", "modality", "weighted_mod"):
        loss = sum(mod_loss.values()) / len(mod_loss)
    elif loss_type == "token":
        loss = sum(mod_loss[m] * mod_count[m] for m in mod_loss) / sum(mod_count.values())
    else:
        raise ValueError("Invalid loss type")
    out["loss"], out["mod_loss"] = loss, mod_loss
    return out
