"""Device-side masking for the mod4 training step.

The reference masks every sample on the CPU inside the data-loader workers (`UnifiedMasking.__call__`,
egom2p/data/masking.py:519-564): a Dirichlet-mixture token budget per modality for the encoder inputs and for the decoder
targets (`input_token_budget` / `target_token_budget`, :181-234) and a random input / target split of each modality's
positions (`image_mask`, :236-266); the (B, 10300) boolean masks are then collated and copied to the GPU. At ~240 samples/s
per GPU that is the step before the hot path that stops keeping up. `DeviceUnifiedMasking` does both on the GPU for a whole
batch: budgets with a handful of vectorised torch ops on (B, n_mod) tensors (torch's CUDA Dirichlet sampler), the masks with
egom2p_image_masks (one CTA per sample and modality: Philox keys, bitonic sort in shared memory). Only token ids cross PCIe.

Same distribution and the same arithmetic as the reference (floor of dirichlet * n, the remainder handed out by argmax of
extra Dirichlet draws, clamp to max_tokens / to what the input left over); given the same noise the masks are bit-identical
to the reference's (tests/test_masking_gpu.py, tests/golden/masking_ref.npz). What cannot be identical is the random stream
itself (CPU torch / numpy / random generators in the workers vs the device Philox stream).
Only the image-like token modalities ('img', 'cam', 'gaze', 'keypoints') -- the whole mod4 set."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from .ops import _p, _s


def image_masks(B: int, L: int, input_budget: torch.Tensor, target_budget: Optional[torch.Tensor], noise: Optional[torch.Tensor] = None,
                stream_id: int = 0, generator: Optional[torch.Generator] = None):
    """input / target / decoder_attention masks (B, L) of one modality. noise (B, L) fp32 >= 0 replaces the Philox draws."""
    lib = _lib.load()
    dev = input_budget.device
    im = torch.empty(B, L, dtype=torch.bool, device=dev)
    tm = torch.empty(B, L, dtype=torch.bool, device=dev)
    cnt = torch.empty(B, L, dtype=torch.int32, device=dev)
    seed = offset = 0
    if noise is None:
        gen = generator if generator is not None else torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        seed, offset = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, gen.get_offset()
        gen.set_offset(offset + 64)          # each thread consumes at most 8 32-bit draws
    else:
        noise = noise.to(torch.float32).contiguous()
    ib = input_budget.to(torch.int32).contiguous()
    tb = target_budget.to(torch.int32).contiguous() if target_budget is not None else None
    _lib.check(lib.egom2p_image_masks(_p(noise), C.c_uint64(seed), C.c_uint64(offset), stream_id, B, L, _p(ib), _p(tb), _p(im), _p(tm),
                                      _p(cnt), _s()), "image_masks")
    return im, tm, cnt


class DeviceUnifiedMasking:
    """Batched, on-device counterpart of egom2p.data.masking.UnifiedMasking for token modalities."""

    def __init__(self, modality_info: Dict, input_tokens_range: Union[int, Tuple[int, int]],
                 target_tokens_range: Optional[Union[int, Tuple[int, int]]], sampling_weights: Optional[Sequence[float]] = None,
                 device: Union[str, torch.device] = "cuda", resample_rounds: int = 8):
        two = lambda r: (r, r) if isinstance(r, int) else tuple(r)
        self.input_tokens_range = two(input_tokens_range)
        self.target_tokens_range = two(target_tokens_range) if target_tokens_range is not None else None
        self.mods: List[str] = list(modality_info)
        for m, inf in modality_info.items():
            if inf["type"] not in ("img", "cam", "gaze", "keypoints"):
                raise NotImplementedError("DeviceUnifiedMasking handles the image-like token modalities (the mod4 set)")
        self.device = torch.device(device)
        t = lambda key, dt: torch.tensor([inf.get(key, 0) for inf in modality_info.values()], dtype=dt, device=self.device)
        self.min_tokens, self.max_tokens = t("min_tokens", torch.int32), t("max_tokens", torch.int32)
        eps = 1e-9
        self.input_alphas = torch.tensor([inf["input_alphas"] for inf in modality_info.values()], dtype=torch.float32,
                                         device=self.device).t().contiguous().clamp(min=eps)      # (n_mix, n_mod)
        self.target_alphas = torch.tensor([inf["target_alphas"] for inf in modality_info.values()], dtype=torch.float32,
                                          device=self.device).t().contiguous().clamp(min=eps)
        self.num_dirichlets = self.input_alphas.shape[0]
        self.sampling_weights = (torch.tensor(list(sampling_weights), dtype=torch.float32, device=self.device)
                                 if sampling_weights is not None else None)
        self.resample_rounds = resample_rounds if bool((self.min_tokens > 0).any()) else 1

    # ---- budgets (masking.py:181-234), vectorised over the batch
    @staticmethod
    def _budget_from_draws(first: torch.Tensor, extra: torch.Tensor, n: torch.Tensor, cap: torch.Tensor) -> torch.Tensor:
        """first (B, n_mod) and extra (B, n_mod, n_mod) Dirichlet draws, n (B,) tokens to hand out, cap (n_mod) or (B, n_mod):
        floor(first * n), the remainder `diff` given to argmax(extra[:, j]) for j < diff, clamped to cap."""
        budget = (first * n[:, None]).floor().to(torch.int32)
        diff = n.to(torch.int32) - budget.sum(1)
        pick = extra.argmax(-1)                                                    # (B, n_mod): modality of the j-th extra token
        use = torch.arange(extra.shape[1], device=first.device)[None, :] < diff[:, None]
        add = torch.zeros_like(budget).scatter_add_(1, pick, use.to(torch.int32))
        return torch.minimum(budget + add, cap.to(torch.int32).expand_as(budget))

    def _sample_budget(self, alphas_b: torch.Tensor, n: torch.Tensor, cap: torch.Tensor) -> torch.Tensor:
        B, nm = alphas_b.shape
        out = None
        for _ in range(self.resample_rounds):
            first = torch._sample_dirichlet(alphas_b)
            extra = torch._sample_dirichlet(alphas_b[:, None, :].expand(B, nm, nm).contiguous())
            cand = self._budget_from_draws(first, extra, n, cap)
            if out is None:
                out = cand
            else:   # keep the first draw that satisfied min_tokens (the reference retries up to max_tries, :196-203)
                ok = (out >= self.min_tokens).all(1, keepdim=True)
                out = torch.where(ok, out, cand)
        return out

    def token_budgets(self, B: int):
        dev = self.device
        if self.sampling_weights is not None:
            dir_idx = torch.multinomial(self.sampling_weights, B, replacement=True)
        else:
            dir_idx = torch.randint(0, self.num_dirichlets, (B,), device=dev)
        lo, hi = self.input_tokens_range
        n_in = torch.randint(lo, hi + 1, (B,), device=dev)
        ib = self._sample_budget(self.input_alphas[dir_idx], n_in, self.max_tokens)
        if self.target_tokens_range is None:
            return ib, None
        lo, hi = self.target_tokens_range
        n_tg = torch.randint(lo, hi + 1, (B,), device=dev)
        cap = torch.maximum(self.min_tokens, self.max_tokens - ib)
        tb = self._sample_budget(self.target_alphas[dir_idx], n_tg, cap)
        return ib, tb

    @torch.no_grad()
    def __call__(self, tokens: Dict[str, torch.Tensor], noise: Optional[Dict[str, torch.Tensor]] = None,
                 budgets: Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]] = None):
        """tokens: {modality: ids (B, ...) on the device} -> mod_dict in the reference layout
        {modality: {tensor, input_mask, target_mask, decoder_attention_mask}} ready for EgoM2P.forward."""
        B = next(iter(tokens.values())).shape[0]
        ib, tb = budgets if budgets is not None else self.token_budgets(B)
        out = {}
        for j, m in enumerate(self.mods):
            L = int(self.max_tokens[j]) if False else tokens[m][0].numel()
            im, tm, cnt = image_masks(B, L, ib[:, j], tb[:, j] if tb is not None else None,
                                      noise=noise[m] if noise is not None else None, stream_id=j)
            out[m] = {"tensor": tokens[m], "input_mask": im, "target_mask": tm, "decoder_attention_mask": cnt}
        return out
