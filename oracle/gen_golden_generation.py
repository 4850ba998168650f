"""TEST INFRASTRUCTURE ONLY. Full-size fixtures for the generation passes (BASELINE.json configs[2..3]): the
UNMODIFIED reference GenerationSampler.forward_enc_dec_roar_batched (egom2p/models/generate.py:747-766) on ego-b weights
(oracle/synth seed 0), fp32 on CPU, conditional and unconditional branch of one guided ROAR step, at the real sizes:

  depth_last : rgb -> depth, third ROAR step: encoder N = 5120 + 3414 = 8534 (cond) / 3414 (uncond), decoder k = 1706
  cam_first  : rgb -> cam, first step: encoder N = 5120 (cond) / 0 (uncond: the empty-context pass), decoder k = 10
  cam_second : rgb -> cam, second step: encoder N = 5130 (cond) / 10 (uncond), decoder k = 10
  rgb_last   : depth -> rgb (BASELINE configs[4]), sixth step: encoder N = 9387 (cond) / 4267 (uncond), decoder k = 853

Stores the decoder positions the reference drew, and per branch the log-sum-exp and a column slice of the logits of every
decoder row, plus argmax / top-2 gap of the guided logits (scale 2.0). Run: `python oracle/gen_golden_generation.py` (~3 min)."""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402
import synth  # noqa: E402
import gen_golden_egob as gg  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "generation_egob.npz")
CASES = {"depth_last": ("tok_depth", 3414, 1706), "cam_first": ("tok_cam", 0, 10), "cam_second": ("tok_cam", 10, 10),
         "rgb_last": ("tok_rgb", 4267, 853)}      # depth -> rgb, sixth ROAR step: N = 5120 + 4267 = 9387 (cond) / 4267 (uncond)
COND = {"depth_last": "tok_rgb", "cam_first": "tok_rgb", "cam_second": "tok_rgb", "rgb_last": "tok_depth"}
COLS = gg.COLS
SCALE = 2.0


def make_state(case: str, device="cpu"):
    """mod_dict of one decoding state: tok_rgb fully given, the target modality with `n_done` tokens already decoded."""
    target, n_done, _ = CASES[case]
    rng = np.random.default_rng(sum(map(ord, case)))
    L = 5120 if target in ("tok_depth", "tok_rgb") else 30
    V = 64000 if L == 5120 else 256
    md = {COND[case]: {"tensor": torch.from_numpy(rng.integers(0, 64000, size=(1, 5, 32, 32), dtype=np.int64)),
                       "input_mask": torch.zeros(1, 5120, dtype=torch.bool), "target_mask": torch.ones(1, 5120, dtype=torch.bool)}}
    ids = np.zeros((1, L), dtype=np.int64)
    im, tm = np.ones((1, L), dtype=bool), np.zeros((1, L), dtype=bool)
    done = rng.permutation(L)[:n_done]
    ids[0, done] = rng.integers(0, V, size=n_done)
    im[0, done], tm[0, done] = False, True
    md[target] = {"tensor": torch.from_numpy(ids), "input_mask": torch.from_numpy(im), "target_mask": torch.from_numpy(tm)}
    return {m: {k: v.to(device) for k, v in d.items()} for m, d in md.items()}


def main():
    import_reference()
    from egom2p.data.modality_info import MODALITY_INFO
    from egom2p.models.generate import GenerationSampler, empty_img_modality
    from egom2p.utils.timm.model_builder import create_model
    torch.set_num_threads(os.cpu_count() or 8)
    torch.set_grad_enabled(False)
    mods = gg.MODS
    model = create_model("egom2p_base_12e_12d_swiglu_nobias",
                         encoder_embeddings={m: MODALITY_INFO[m]["encoder_embedding"]() for m in mods},
                         decoder_embeddings={m: MODALITY_INFO[m]["decoder_embedding"]() for m in mods},
                         modality_info={m: MODALITY_INFO[m] for m in mods}, num_register_tokens=0).eval()
    model.load_state_dict(synth.make_state_dict(gg.egob_cfg(), gg.SD_SEED), strict=True)
    sampler = GenerationSampler(model)
    res = {"cols": COLS}
    only = sys.argv[1:]
    if only and os.path.exists(OUT):   # add / refresh single cases without recomputing the others
        res.update({k: v for k, v in np.load(OUT).items()})
    for case, (target, n_done, k) in CASES.items():
        if only and case not in only:
            continue
        md = make_state(case)
        lc, pos = sampler.forward_enc_dec_roar_batched(copy.deepcopy(md), target, k, seed=5)
        mu = empty_img_modality(copy.deepcopy(md), COND[case])
        lu, pos_u = sampler.forward_enc_dec_roar_batched(mu, target, k, seed=5)
        assert torch.equal(pos, pos_u)
        res[f"{case}::pos"] = pos.numpy()
        for name, lg in (("cond", lc), ("uncond", lu)):
            lg = lg[0]
            res[f"{case}::{name}::lse"] = torch.logsumexp(lg.double(), -1).numpy()
            res[f"{case}::{name}::logits"] = lg.numpy().copy() if lg.shape[-1] <= 256 else lg[:, torch.from_numpy(COLS)].numpy().copy()
        guided = (lu + (lc - lu) * SCALE)[0].double()
        top2 = torch.topk(guided, 2, dim=-1)
        res[f"{case}::guided_argmax"] = top2[1][:, 0].numpy()
        res[f"{case}::guided_gap"] = (top2[0][:, 0] - top2[0][:, 1]).numpy()
        print(case, "pos", tuple(pos.shape), "lse cond", float(res[f"{case}::cond::lse"].mean()), "gap median", float(np.median(res[f"{case}::guided_gap"])), flush=True)
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
