"""Generation with the B200 model: guided ROAR / MaskGIT decoding of token modalities.

Mirrors the reference's generation API for the image-like token modalities the mod4 model has ('img', 'cam', 'gaze'):
`GenerationSampler(model).generate(mod_dict, schedule, top_k, top_p, seed=...)` with the schedule dictionaries of
`build_chained_generation_schedules`, and the mod_dict helpers `init_empty_target_modality` / `init_full_input_modality` /
`empty_img_modality` (reference: egom2p/models/generate.py:30-37,83-152,197-321,323-405,747-817,1031-1099;
egom2p/utils/generation.py:49-105). The eval scripts (eval_model_rgb2depth.py:45-95 and siblings) run unchanged against it.

What is different underneath (same results, B200-first mechanics):
  * one decoding step = encoder pass(es) straight from token ids through the index plan + fused embed kernel (no
    (B, L, D) embedding of whole modalities, no argsort / gathers), one decoder pass, one head GEMM, one sampling kernel;
  * classifier-free guidance: the conditional and the unconditional branch share ONE decoder batch (their contexts differ
    in length; a branch with an empty context contributes exactly 0 in cross-attention), and the guidance combine
    l_u + s (l_c - l_u) is applied to the decoder outputs before the bias-free linear head, so a single head GEMM produces
    the guided logits; nothing is deep-copied per step (the reference copies the whole mod_dict, generate.py:793);
  * sampling (temperature, top-k, top-p, categorical draw) runs in egom2p_sample_rows on row chunks sized to stay in L2:
    the (B, k, 64000) fp32 logits tensor (three of them in the reference) is never materialised, nothing is sorted;
  * token counts are tracked on the host from the schedule: no `.max()` / `.sum()` device-to-host syncs per step (the
    reference syncs in forward_mask_encoder_generation, generate.py:415, and forward_mask_decoder_roar, :495).
Random positions of ROAR are drawn exactly as the reference does (torch.manual_seed(seed + step); torch.rand(L) * 1e-6 added
to the target mask; argsort), so the decoding ORDER matches the reference for the same seed; the categorical draws use the
device generator through torch.rand (torch.multinomial's internal stream cannot be reproduced), which only matters at
temperatures where the reference itself is stochastic (the eval scripts decode at T = 0.01)."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import nn

from . import ops

f32 = torch.float32


# ----------------------------------------------------------------------------------------------- mod_dict helpers
def empty_img_modality(mod_dict, key):
    mod_dict[key]["input_mask"][:] = True
    mod_dict[key]["target_mask"][:] = False
    return mod_dict


def init_empty_target_modality(mod_dict, modality_info, domain, batch_size, num_tokens, device):
    if modality_info[domain]["type"] not in ("img", "gaze", "cam", "keypoints"):
        raise NotImplementedError("egom2p_b200 generates token modalities of type img / cam / gaze (the mod4 set)")
    mod_dict[domain] = {"tensor": torch.zeros((batch_size, num_tokens), dtype=torch.int64, device=device),
                        "input_mask": torch.ones((batch_size, num_tokens), dtype=torch.bool, device=device),
                        "target_mask": torch.zeros((batch_size, num_tokens), dtype=torch.bool, device=device)}
    return empty_img_modality(mod_dict, domain)


def init_full_input_modality(mod_dict, modality_info, domain, device, eos_id=3):
    if modality_info[domain]["type"] not in ("img", "gaze", "cam", "keypoints"):
        raise NotImplementedError("egom2p_b200 generates token modalities of type img / cam / gaze (the mod4 set)")
    shape = mod_dict[domain]["tensor"].shape
    shape = (shape[0], int(np.prod(shape[1:])))
    for key, fill in (("input_mask", False), ("target_mask", True), ("decoder_attention_mask", False)):
        if key not in mod_dict[domain]:
            mod_dict[domain][key] = torch.full(shape, fill, dtype=torch.bool, device=device)
    mod_dict[domain]["input_mask"][:] = False
    mod_dict[domain]["target_mask"][:] = True
    return mod_dict


# ----------------------------------------------------------------------------------------------- schedules
def linear_schedule(num_steps, total_tokens):
    schedule = np.linspace(0, total_tokens, num_steps + 1, dtype=int)
    tokens = np.sort(np.diff(schedule))[::-1]
    return np.trim_zeros(tokens, "b")


def cosine_schedule(num_steps, total_tokens):
    vals = np.array([0.5 * (1 + math.cos(math.pi * i / num_steps)) for i in range(num_steps)])
    tokens = [round(total_tokens * d) for d in (vals[:-1] - vals[1:])]
    tokens.append(total_tokens - sum(tokens))
    return np.array(tokens)


def linear_temp_schedule(temp, token_schedule):
    total = token_schedule.sum()
    return np.concatenate([np.array([temp * 1.0]), (temp * (total - token_schedule.cumsum()) / total)[:-1]]).clip(min=1e-9)


def onex_temp_schedule(max_t, min_t, token_schedule, power=0.5, min_linspace=1, max_linspace=100):
    x = np.linspace(min_linspace, max_linspace, num=sum(token_schedule))
    y = 1 / (x ** power)
    y = y - min(y)
    y = y / max(y)
    frac = np.cumsum(token_schedule) / np.sum(token_schedule)
    unscaled = [(1 - cs) * us for us, cs in zip(y, frac)]
    return np.array([min_t + (max_t - min_t) * s for s in unscaled]).clip(min=1e-9)


def build_chained_generation_schedules(cond_domains: List[str], target_domains: List[str], tokens_per_target: List[int],
                                       autoregression_schemes: List[str], decoding_steps: List[int],
                                       token_decoding_schedules: List[str], temps: List[float], temp_schedules: List[str],
                                       cfg_scales: List[float], cfg_schedules: List[str], cfg_grow_conditioning: bool = False,
                                       modality_info: Optional[dict] = None):
    """List of {target_domain, scheme, num_tokens, temperature, cfg_scale, cfg_cond_domains} steps (generate.py:197-321)."""
    chained = []
    cond_domains = list(cond_domains)
    for i, target in enumerate(target_domains):
        scheme, ntoks, steps = autoregression_schemes[i], tokens_per_target[i], decoding_steps[i]
        if scheme == "autoregressive":
            raise NotImplementedError("autoregressive (sequence) targets are not part of the mod4 path")
        if scheme == "maskgit":
            if token_decoding_schedules[i] == "cosine":
                tok = cosine_schedule(steps, ntoks)
            elif token_decoding_schedules[i] == "linear":
                tok = linear_schedule(steps, ntoks)
            else:
                raise ValueError(f"Illegal MaskGIT token schedule {token_decoding_schedules[i]}")
        elif scheme == "roar":
            tok = linear_schedule(steps, ntoks)
        else:
            raise ValueError(f"Illegal decoding scheme {scheme}")
        name = temp_schedules[i]
        if name == "linear":
            tsched = linear_temp_schedule(temps[i], tok)
        elif name == "constant":
            tsched = temps[i] * np.ones(steps)
        elif "onex" in name:
            min_t, power = [float(f) for f in name.split(":")[1:]]
            tsched = onex_temp_schedule(max_t=temps[i], min_t=min_t, token_schedule=tok, power=power)
        else:
            raise ValueError(f"Illegal temperature schedule {name}")
        if cfg_schedules[i] != "constant":
            raise NotImplementedError() if cfg_schedules[i] == "cosine" else ValueError(f"Illegal guidance schedule {cfg_schedules[i]}")
        cfg = cfg_scales[i] * np.ones(steps)
        chained.extend({"target_domain": target, "scheme": scheme, "num_tokens": k, "temperature": t, "cfg_scale": c,
                        "cfg_cond_domains": list(cond_domains)} for k, t, c in zip(tok, tsched, cfg))
        if cfg_grow_conditioning:
            cond_domains.append(target)
    return chained


# ----------------------------------------------------------------------------------------------- the sampler
class GenerationSampler(nn.Module):
    """Drop-in for egom2p.models.generate.GenerationSampler on the token modalities of the mod4 model."""
    LOGIT_CHUNK_BYTES = 48 << 20   # fp32 logits of one head-GEMM + sampling chunk: stays in the 126 MB L2

    def __init__(self, model):
        super().__init__()
        self.model = model
        self.stats = {"steps": 0, "encoder_passes": 0, "decoder_rows": 0, "head_rows": 0}

    # -- sampling of a block of decoder outputs (f1)
    def sample_from_hidden(self, yb: torch.Tensor, target_mod: str, temperature: float, top_k=0.0, top_p=0.0):
        """yb (rows, D) bf16 decoder outputs (guidance already applied) -> (tokens int64, probs fp32), chunk by chunk through
        head GEMM + egom2p_sample_rows. Same filtering semantics as the reference's sample_tokens (generate.py:361-371)."""
        wb = self.model.head_operand(target_mod)
        V = wb.shape[0]
        if isinstance(top_k, float) and top_k > 0.0:
            top_k = min(int(top_k * V), V)
        rows = yb.shape[0]
        chunk = max(8, min(rows, self.LOGIT_CHUNK_BYTES // (4 * V)))
        toks, probs = [], []
        greedy = bool(np.isclose(temperature, 0, atol=1e-10))
        u_all = torch.rand(rows, device=yb.device, dtype=f32)
        for r0 in range(0, rows, chunk):
            lg = ops.linear_fwd(yb[r0:r0 + chunk], wb, out_dtype=f32)
            t, p, _ = ops.sample_rows(lg, 0.0 if greedy else float(temperature), float(top_p), int(top_k), u_all[r0:r0 + chunk])
            toks.append(t); probs.append(p)
        self.stats["head_rows"] += rows
        return torch.cat(toks), torch.cat(probs)

    # -- decoder positions of one step
    def _positions(self, target_mask: torch.Tensor, n_left: int, scheme: str, num_select: int, seed: Optional[int]):
        """(B, L) target mask -> (B, k) positions of the decoder tokens of this step, as the reference picks them:
        ROAR: a random subset of the open positions, the same permutation for every sample (generate.py:481-516);
        MaskGIT: all open positions in ascending order (:447-479)."""
        L = target_mask.shape[1]
        if seed is not None:
            torch.manual_seed(seed)
        if scheme == "roar":
            k = min(int(num_select), n_left)
            noise = torch.rand(L, device=target_mask.device).unsqueeze(0) * 1e-6
        else:
            k = n_left
            noise = torch.arange(L, device=target_mask.device).unsqueeze(0) * 1e-6
        return torch.argsort(target_mask + noise, dim=1)[:, :k]

    def _decoder_inputs(self, target_mod: str, pos: torch.Tensor, target_mask: torch.Tensor):
        """mask_token + (pos_emb[pos] + mod_emb) for the selected positions through the fused embed kernel (decoder mode)."""
        m = self.model
        emb = m.decoder_embeddings[target_mod]
        B, k = pos.shape
        plan = ops.Plan()
        plan.B, plan.budget = B, k
        plan.keep_mod = torch.zeros(B, k, dtype=torch.int32, device=pos.device)
        plan.keep_pos = pos.to(torch.int32).contiguous()
        plan.pad = torch.gather(target_mask, 1, pos).contiguous()      # open positions are valid (False)
        L = emb.pos_emb.shape[1]
        y0, _ = ops.embed_gather_fwd(plan, m.dim, [L], [int(emb.vocab_size)], None, None, [emb.pos_emb.detach().reshape(-1, m.dim)],
                                     [emb.mod_emb.detach().reshape(-1)], mask_token=m.mask_token.detach().reshape(-1), want_emb=False)
        return y0

    def count_tokens(self, mod_dict):
        """[[valid inputs per sample], [open targets per sample]] per encoder modality, as host lists (one sync)."""
        enc_mods = [mod for mod in mod_dict if mod in self.model.encoder_embeddings]
        B = mod_dict[enc_mods[0]]["tensor"].shape[0]
        flat = lambda t: t.reshape(B, -1)
        return torch.stack([torch.stack([(~flat(mod_dict[mod]["input_mask"])).sum(1), (~flat(mod_dict[mod]["target_mask"])).sum(1)])
                            for mod in enc_mods]).tolist()

    @torch.no_grad()
    def forward_hidden(self, mod_dict, target: str, uncond_without: List[str], n_in: Dict[str, List[int]], pos: torch.Tensor):
        """Decoder outputs (after decoder_norm) of one decoding step for the conditional branch and -- if `uncond_without`
        names conditioning modalities -- the unconditional branch, as (branches, B * k, D) fp32 with branch 0 = conditional.
        The reference runs forward_enc_dec_roar_batched twice on two copies of the mod_dict (generate.py:789-802); here the
        encoder runs once per branch at that branch's own length (not at all for an empty one) and the decoder ONCE over both
        branches. n_in: valid input tokens per modality and sample (host ints); pos (B, k): decoder positions in `target`."""
        m = self.model
        enc_mods = [mod for mod in mod_dict if mod in m.encoder_embeddings]
        B, k = pos.shape
        dev = pos.device
        flat = lambda t: t.reshape(B, -1)
        masks_c = {mod: flat(mod_dict[mod]["input_mask"]) for mod in enc_mods}
        branches = [(masks_c, [sum(n_in[mod][b] for mod in enc_mods) for b in range(B)])]
        if uncond_without:
            masks_u = {mod: (torch.ones_like(masks_c[mod]) if mod in uncond_without else masks_c[mod]) for mod in enc_mods}
            branches.append((masks_u, [sum(n_in[mod][b] for mod in enc_mods if mod not in uncond_without) for b in range(B)]))
        ctxs = []
        for masks, tot in branches:
            if max(tot) == 0:
                ctxs.append((None, None))
                continue
            ctxs.append(m.encode_tokens(mod_dict, masks, max(tot)))
            self.stats["encoder_passes"] += 1
        y0 = self._decoder_inputs(target, pos, flat(mod_dict[target]["target_mask"]))
        nb = len(branches)
        n_max = max((c.shape[1] for c, _ in ctxs if c is not None), default=0)
        if n_max == 0:
            yn = m.decode_tokens(y0.repeat(nb, 1, 1) if nb > 1 else y0, None, None)
        else:
            parts, lens = [], []
            for c, nv in ctxs:
                if c is None:
                    parts.append(torch.zeros(B, n_max, m.dim, dtype=f32, device=dev))
                    lens.append(torch.zeros(B, dtype=torch.int32, device=dev))
                else:
                    parts.append(c if c.shape[1] == n_max else torch.nn.functional.pad(c, (0, 0, 0, n_max - c.shape[1])))
                    lens.append(nv)
            yn = m.decode_tokens(y0.repeat(nb, 1, 1) if nb > 1 else y0, torch.cat(parts) if nb > 1 else parts[0],
                                 torch.cat(lens) if nb > 1 else lens[0])
        self.stats["decoder_rows"] += nb * B * k
        return yn.reshape(nb, B * k, m.dim)

    @torch.no_grad()
    def generate(self, mod_dict, schedule, top_k=0.0, top_p=0.0, text_tokenizer=None, verbose=False, seed=None, _counts=None):
        m = self.model
        info = m.modality_info
        mod_dict = {mod: {k: (v.clone() if k in ("tensor", "input_mask", "target_mask") else v) for k, v in d.items()}
                    for mod, d in mod_dict.items()}
        enc_mods = [mod for mod in mod_dict if mod in m.encoder_embeddings]
        B = mod_dict[enc_mods[0]]["tensor"].shape[0]
        dev = mod_dict[enc_mods[0]]["tensor"].device
        flat = lambda t: t.reshape(B, -1)
        # the only device -> host transfer of the whole generation: valid input / open target counts per modality and sample
        # (`_counts`: supplied by GraphedGeneration, which captures this method in a CUDA graph)
        cnt = _counts if _counts is not None else self.count_tokens(mod_dict)
        n_in = {mod: list(map(int, cnt[i][0])) for i, mod in enumerate(enc_mods)}
        n_open = {mod: int(cnt[i][1][0]) for i, mod in enumerate(enc_mods)}   # the reference assumes equal counts in a batch (:462)

        for step, st in enumerate(schedule):
            target, temp = st["target_domain"], float(st["temperature"])
            scale, cond = float(st.get("cfg_scale", 1.0)), list(st.get("cfg_cond_domains", []))
            if info[target]["type"] not in ("img", "cam", "gaze", "keypoints"):
                raise NotImplementedError("egom2p_b200 generates token modalities of type img / cam / gaze (the mod4 set)")
            scheme = st["scheme"].lower()
            if scheme not in ("roar", "maskgit"):
                raise ValueError("Invalid sampling scheme")
            seed_i = seed + step if seed is not None else None
            guided = not (scale == 1.0 or len(cond) == 0)

            tmask = flat(mod_dict[target]["target_mask"])
            pos = self._positions(tmask, n_open[target], scheme, st["num_tokens"], seed_i)
            k = pos.shape[1]
            if k == 0:
                continue
            yn = self.forward_hidden(mod_dict, target, cond if guided else [], n_in, pos)
            yb = ops.cfg_combine_bf16(yn[1].contiguous(), yn[0].contiguous(), scale) if guided else ops.cast_bf16(yn[0].contiguous())

            # ---- head + sampling (fused, chunked), then write the tokens back
            samples, probs = self.sample_from_hidden(yb, target, temp, top_k, top_p)
            samples, probs = samples.reshape(B, k), probs.reshape(B, k)
            if scheme == "maskgit":
                n_sel = min(int(st["num_tokens"]), k)
                top = torch.topk(probs, n_sel, dim=-1)[1]
                pos, samples = torch.gather(pos, -1, top), torch.gather(samples, -1, top)
                k = n_sel
            d = mod_dict[target]
            d["tensor"] = torch.scatter(flat(d["tensor"]), -1, pos, samples.to(d["tensor"].dtype)).reshape(d["tensor"].shape)
            d["input_mask"] = torch.scatter(flat(d["input_mask"]), -1, pos, torch.zeros_like(samples, dtype=torch.bool)).reshape(d["input_mask"].shape)
            d["target_mask"] = torch.scatter(flat(d["target_mask"]), -1, pos, torch.ones_like(samples, dtype=torch.bool)).reshape(d["target_mask"].shape)
            n_in[target] = [v + k for v in n_in[target]]
            n_open[target] -= k
            self.stats["steps"] += 1
        return mod_dict


class GraphedGeneration:
    """One CUDA graph for a whole `GenerationSampler.generate` call of fixed shapes (workload, batch, schedule).

    At batch 1 a guided 3-step decode is ~1100 kernel launches of a few microseconds each: issued from Python the call is
    bound by the host (rgb -> cam: 24.6 ms per clip eager). Every shape in `generate` follows from the schedule and the
    token counts of the initial mod_dict, so the whole call -- encoder passes, batched decoder pass, head GEMM, sampling
    kernel and the scatter of the new tokens, for all steps -- is captured once and replayed: new conditioning tokens are
    copied into static buffers, the generated tokens are read from static outputs. The ROAR positions and the categorical
    draws use torch's graph-safe Philox generator (each replay draws fresh numbers); the per-step `torch.manual_seed` of the
    eager path is not replayable and is skipped (seed=None)."""

    def __init__(self, sampler: GenerationSampler, example_mod_dict, schedule, top_k=0.0, top_p=0.0, warmup: int = 2):
        self.sampler, self.schedule, self.top_k, self.top_p = sampler, schedule, top_k, top_p
        self.static_in = {mod: {k: v.clone() for k, v in d.items()} for mod, d in example_mod_dict.items()}
        self.counts = sampler.count_tokens(self.static_in)
        dev = next(iter(self.static_in.values()))["tensor"].device
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = self._run()
        torch.cuda.synchronize(dev)

    def _run(self):
        return self.sampler.generate(self.static_in, self.schedule, top_k=self.top_k, top_p=self.top_p, seed=None, _counts=self.counts)

    @torch.no_grad()
    def __call__(self, mod_dict):
        """mod_dict with the same shapes and the same input / target mask pattern as the example; returns the generated
        mod_dict (static tensors, overwritten by the next call)."""
        for mod, d in self.static_in.items():
            for k, v in d.items():
                v.copy_(mod_dict[mod][k], non_blocking=True)
        self.graph.replay()
        return self.static_out
