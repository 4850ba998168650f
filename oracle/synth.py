"""TEST INFRASTRUCTURE ONLY. Deterministic synthetic weights / batches shared by the golden generator
(which loads them into the live reference) and the tests (which feed them to the oracle and to the
CUDA path). numpy PCG64 streams are platform-stable, so fixtures only need to store outputs."""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

from egom2p_oracle import sincos_1d, sincos_3d

# modality ids = sha256(name) % 2**15 (egom2p/utils/misc.py:39-41, egom2p/data/modality_info.py)
MOD_IDS = {"tok_cam": 349, "tok_depth": 6323, "tok_gaze": 26680, "tok_rgb": 7613}


def make_cfg(dim, heads, enc_depth, dec_depth, mods: List[str], video_vocab=64000, video_thw=(5, 32, 32),
             small_vocab=256, small_len=30):
    info = {}
    for m in mods:
        if m in ("tok_rgb", "tok_depth"):
            info[m] = {"id": MOD_IDS[m], "vocab": video_vocab, "thw": tuple(video_thw),
                       "len": int(np.prod(video_thw)), "type": "img"}
        else:
            info[m] = {"id": MOD_IDS[m], "vocab": small_vocab, "len": small_len,
                       "type": m.split("_")[1]}
    return {"dim": dim, "heads": heads, "enc_depth": enc_depth, "dec_depth": dec_depth, "mods": info,
            "hidden": int(2 * (dim * 4) / 3)}


def state_dict_manifest(cfg) -> Dict[str, tuple]:
    """Key -> shape for the swiglu/no-bias family, in the reference's state_dict layout
    (SURVEY.md A10): parameters, tied duplicates and persistent buffers."""
    D, F = cfg["dim"], cfg["hidden"]
    man: Dict[str, tuple] = {}
    for side in ("encoder_embeddings", "decoder_embeddings"):
        for m, inf in cfg["mods"].items():
            p = f"{side}.{m}."
            man[p + "pos_emb"] = (1, inf["len"], D)
            man[p + "mod_emb"] = (1, 1, D)
            man[p + "token_emb.weight"] = (inf["vocab"], D)
            if side == "decoder_embeddings":
                man[p + "to_logits.weight"] = (inf["vocab"], D)
    def norm(p):
        man[p + "weight"] = (D,)
        man[p + "bias"] = (D,)
    def mlp(p):
        man[p + "fc1.weight"] = (F, D)
        man[p + "fc2.weight"] = (D, F)
        man[p + "fc3.weight"] = (F, D)
    for i in range(cfg["enc_depth"]):
        p = f"encoder.{i}."
        norm(p + "norm1.")
        man[p + "attn.qkv.weight"] = (3 * D, D)
        man[p + "attn.proj.weight"] = (D, D)
        norm(p + "norm2.")
        mlp(p + "mlp.")
    norm("encoder_norm.")
    man["decoder_proj_context.weight"] = (D, D)
    man["decoder_proj_context.bias"] = (D,)
    for i in range(cfg["dec_depth"]):
        p = f"decoder.{i}."
        norm(p + "norm1.")
        man[p + "self_attn.qkv.weight"] = (3 * D, D)
        man[p + "self_attn.proj.weight"] = (D, D)
        man[p + "cross_attn.q.weight"] = (D, D)
        man[p + "cross_attn.kv.weight"] = (2 * D, D)
        man[p + "cross_attn.proj.weight"] = (D, D)
        norm(p + "query_norm.")
        norm(p + "context_norm.")
        norm(p + "norm2.")
        mlp(p + "mlp.")
    norm("decoder_norm.")
    man["mask_token"] = (1, 1, D)
    return man


def make_state_dict(cfg, seed: int, tie: bool = True) -> Dict[str, torch.Tensor]:
    """Random but reference-plausible weights: matrices ~U(+-sqrt(6/(fan_in+fan_out))), embeddings and
    tokens ~N(0, 0.02) (tied heads take the embedding), norm weights 1 + N(0, 0.1) (so LN-weight paths are
    exercised), LN bias buffers 0, context bias N(0, 0.02). pos_emb = the fixed sin-cos tables."""
    rng = np.random.default_rng(seed)
    sd: Dict[str, torch.Tensor] = {}
    for k, shp in state_dict_manifest(cfg).items():
        if k.endswith("pos_emb"):
            m = k.split(".")[1]
            inf = cfg["mods"][m]
            tab = sincos_3d(*inf["thw"], cfg["dim"]) if "thw" in inf else sincos_1d(inf["len"], cfg["dim"])
            sd[k] = tab[None].clone()
        elif k.endswith("to_logits.weight"):
            continue
        elif k.endswith(".bias") and k != "decoder_proj_context.bias":
            sd[k] = torch.zeros(shp)
        elif "norm" in k and k.endswith(".weight"):
            sd[k] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(shp)).astype(np.float32))
        elif len(shp) == 2 and "token_emb" not in k:
            b = math.sqrt(6.0 / (shp[0] + shp[1]))
            sd[k] = torch.from_numpy(rng.uniform(-b, b, shp).astype(np.float32))
        else:
            sd[k] = torch.from_numpy((0.02 * rng.standard_normal(shp)).astype(np.float32))
    for m in cfg["mods"]:
        p = f"decoder_embeddings.{m}."
        sd[p + "mod_emb"] = sd[f"encoder_embeddings.{m}.mod_emb"]  # share_modality_embeddings (egom2p_model.py:179-183)
        if tie:
            sd[p + "to_logits.weight"] = sd[p + "token_emb.weight"]
        else:
            shp = sd[p + "token_emb.weight"].shape
            b = math.sqrt(6.0 / (shp[0] + shp[1]))
            sd[p + "to_logits.weight"] = torch.from_numpy(rng.uniform(-b, b, tuple(shp)).astype(np.float32))
    return sd


def make_batch(cfg, B: int, seed: int, n_in: Dict[str, List[int]], n_tgt: Dict[str, List[int]]):
    """mod_dict in the reference's layout (egom2p/data/masking.py:236-266): per modality a random
    permutation; the first n_in positions of it are inputs, the next n_tgt are targets; the
    decoder_attention_mask holds the target count at the first target position."""
    rng = np.random.default_rng(seed)
    md = {}
    for m, inf in cfg["mods"].items():
        L = inf["len"]
        ids = rng.integers(0, inf["vocab"], size=(B, L), dtype=np.int64)
        imask = np.ones((B, L), dtype=bool)
        tmask = np.ones((B, L), dtype=bool)
        cnt = np.zeros((B, L), dtype=np.int32)
        for b in range(B):
            perm = rng.permutation(L)
            ni, nt = n_in[m][b], n_tgt[m][b]
            imask[b, perm[:ni]] = False
            tmask[b, perm[ni:ni + nt]] = False
            first = int(np.argmin(tmask[b].astype(np.float32) + np.arange(L) * 1e-6))
            cnt[b, first] = nt
        t = torch.from_numpy(ids)
        if "thw" in inf:
            t = t.reshape(B, *inf["thw"])
        md[m] = {"tensor": t, "input_mask": torch.from_numpy(imask), "target_mask": torch.from_numpy(tmask),
                 "decoder_attention_mask": torch.from_numpy(cnt)}
    return md
