"""TEST INFRASTRUCTURE ONLY. Pins the device-side masking against the UNMODIFIED reference UnifiedMasking
(egom2p/data/masking.py, imported from /root/reference): (1) image_mask for given budgets, with the noise torch.rand drew
stored beside the masks it produced; (2) the budget arithmetic of input_token_budget / target_token_budget with the
Dirichlet objects replaced by a replay of stored draws, so that the vectorised restatement can be checked draw for draw.
-> tests/golden/masking_ref.npz. Run: `python oracle/gen_golden_masking.py`."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402
from gen_golden_masks import _Tok, MODS, ALPHAS  # noqa: E402  (imports the reference too)

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "masking_ref.npz")


class _Replay:
    """Stands in for a torch Dirichlet: hands out stored draws in order (sample(): one; sample_n(k): the next k)."""

    def __init__(self, draws):
        self.draws, self.i = draws, 0

    def sample(self):
        self.i += 1
        return self.draws[self.i - 1]

    def sample_n(self, k):
        k = int(k)
        out = self.draws[self.i:self.i + k] if k > 0 else self.draws[:0]
        self.i += k
        return out


def main():
    import_reference()
    from egom2p.data.masking import UnifiedMasking
    from egom2p.data.modality_info import MODALITY_INFO
    info = {}
    for m in MODS:
        d = dict(MODALITY_INFO[m])
        d["input_alphas"], d["target_alphas"] = ALPHAS, ALPHAS
        d.setdefault("min_tokens", 0)
        info[m] = d
    masking = UnifiedMasking(modality_info=info, text_tokenizer=_Tok(), input_tokens_range=(2048, 2048),
                             target_tokens_range=(2048, 2048), sampling_weights=[1.0, 1.0, 1.0, 1.0])
    res = {}
    # ---- (1) image_mask: noise + masks for a few (L, input budget, target budget) cases
    cases = [(5120, 1009, 1009), (5120, 0, 2048), (5120, 5120, 0), (5120, 37, 5000), (30, 15, 15), (30, 0, 0), (30, 30, 0), (30, 7, 23)]
    for i, (L, ib, tb) in enumerate(cases):
        torch.manual_seed(100 + i)
        noise = torch.rand(L)
        torch.manual_seed(100 + i)
        out = masking.image_mask(torch.zeros(L, dtype=torch.int64), L, ib, tb)
        res[f"im{i}::cfg"] = np.array([L, ib, tb])
        res[f"im{i}::noise"] = noise.numpy()
        res[f"im{i}::input_mask"] = out["input_mask"].numpy()
        res[f"im{i}::target_mask"] = out["target_mask"].numpy()
        res[f"im{i}::attn"] = out["decoder_attention_mask"].numpy()
    res["n_image_cases"] = np.array(len(cases))
    # ---- (2) budgets from replayed Dirichlet draws (first draw + up to n_mod extra draws per budget)
    g = torch.Generator().manual_seed(7)
    n = 48
    nm = len(MODS)
    ins, tgs, d_in, d_tg, mix = [], [], [], [], []
    for s in range(n):
        k = s % len(ALPHAS)
        alpha = torch.tensor([ALPHAS[k]] * nm).clamp(min=1e-9)
        draws_in = torch._sample_dirichlet(alpha[None].expand(1 + nm, nm).contiguous(), g)
        draws_tg = torch._sample_dirichlet(alpha[None].expand(1 + nm, nm).contiguous(), g)
        masking.input_dirichlets[k] = _Replay(draws_in)
        masking.target_dirichlets[k] = _Replay(draws_tg)
        ib = masking.input_token_budget(2048, k)
        tb = masking.target_token_budget(ib, 2048, k)
        ins.append(ib); tgs.append(tb); d_in.append(draws_in.numpy()); d_tg.append(draws_tg.numpy()); mix.append(k)
    res.update(budget_in=np.array(ins), budget_tg=np.array(tgs), draws_in=np.stack(d_in), draws_tg=np.stack(d_tg), mix=np.array(mix),
               max_tokens=np.array([info[m]["max_tokens"] for m in MODS]))
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT), "first budgets", ins[:4], tgs[:4])


if __name__ == "__main__":
    main()
