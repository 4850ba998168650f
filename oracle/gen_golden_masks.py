"""TEST INFRASTRUCTURE ONLY. Draws masks for the ego-b mod4 step from the UNMODIFIED reference masking code
(egom2p/data/masking.py: UnifiedMasking, input / target budgets 2048 / 2048, the Dirichlet mixture of
cfgs/default/egom2p/alphas_mixture/main/mix_mod4_all2all_uni.yaml, torch / numpy / random seeded with 0) -- the
"reference-distribution" regime of SURVEY.md section 8(d) -- and stores them bit-packed in
tests/golden/ref_masks_egob.npz for `bench.py --regime reference-masks` and tests/test_ragged_masks_gpu.py.
Run: `python oracle/gen_golden_masks.py [n_samples]`."""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

import_reference()
from egom2p.data.masking import UnifiedMasking  # noqa: E402
from egom2p.data.modality_info import MODALITY_INFO  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_masks_egob.npz")
MODS = ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"]  # sorted domains, as run_training_egom2p.py:274-276
ALPHAS = [0.01, 0.1, 1.0, 10.0]


class _Tok:  # the masking class only asks the text tokenizer for sentinel / pad / eos ids (unused by token modalities)
    def get_vocab(self):
        return {"[S_0]": 4, "[S_1]": 5}

    def token_to_id(self, t):
        return {"[PAD]": 0, "[EOS]": 3}.get(t, 1)


def main(n=64):
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    info = {}
    for m in MODS:
        d = dict(MODALITY_INFO[m])
        d["input_alphas"], d["target_alphas"] = ALPHAS, ALPHAS
        d.setdefault("min_tokens", 0)
        info[m] = d
    masking = UnifiedMasking(modality_info=info, text_tokenizer=_Tok(), input_tokens_range=(2048, 2048),
                             target_tokens_range=(2048, 2048), sampling_weights=[1.0, 1.0, 1.0, 1.0])
    store = {}
    stats = []
    for i in range(n):
        sample = {m: torch.zeros(info[m]["max_tokens"], dtype=torch.int64) for m in MODS}
        out = masking(sample)
        nin = ntg = 0
        for m in MODS:
            im, tm = out[m]["input_mask"].numpy().astype(bool), out[m]["target_mask"].numpy().astype(bool)
            store.setdefault(m + "_input_mask", []).append(np.packbits(im))
            store.setdefault(m + "_target_mask", []).append(np.packbits(tm))
            store.setdefault(m + "_attn", []).append(out[m]["decoder_attention_mask"].numpy().astype(np.int32))
            nin += int((~im).sum()); ntg += int((~tm).sum())
        stats.append((nin, ntg))
    np.savez_compressed(OUT, n=np.array(n), valid=np.array(stats, dtype=np.int32),
                        **{k: np.stack(v) for k, v in store.items()},
                        **{m + "_len": np.array(info[m]["max_tokens"]) for m in MODS})
    print("first 8 (valid inputs, valid targets):", stats[:8])
    print("mean valid inputs %.0f, targets %.0f of 2048" % (np.mean([s[0] for s in stats]), np.mean([s[1] for s in stats])))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
