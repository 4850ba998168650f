"""TEST INFRASTRUCTURE ONLY. Runs the UNMODIFIED reference GenerationSampler.top_k_top_p_filtering / sample_tokens
(egom2p/models/generate.py:332-371, imported from /root/reference, authoring container only) on seeded random logits and
stores which tokens survive the filter plus the resulting probabilities of a few tokens -> tests/golden/sampling_filter.npz.
Run: `python oracle/gen_golden_sampling.py`."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "sampling_filter.npz")
CASES = [  # (name, rows, V, logit scale, top_k, top_p, temperature)
    ("cam_p08", 24, 256, 3.0, 0.0, 0.8, 1.0),
    ("vid_p08_t001", 6, 64000, 2.0, 0.0, 0.8, 0.01),
    ("vid_p095", 4, 64000, 4.0, 0.0, 0.95, 1.0),
    ("vid_k50", 4, 64000, 1.0, 50, 0.0, 0.7),
    ("cam_k01_p05", 16, 256, 2.0, 0.1, 0.5, 1.3),
]


def make_logits(name, rows, V, scale):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    return (rng.standard_normal((rows, V)) * scale).astype(np.float32)


def main():
    import_reference()
    from egom2p.models.generate import GenerationSampler
    sampler = GenerationSampler(torch.nn.Identity())
    res = {}
    for name, rows, V, scale, top_k, top_p, temp in CASES:
        lg = make_logits(name, rows, V, scale)
        filt = sampler.top_k_top_p_filtering(torch.from_numpy(lg.copy()), top_k=top_k, top_p=top_p)
        keep = torch.isfinite(filt).numpy()
        pr = torch.softmax(filt / temp, dim=-1).numpy()
        res[name + "::keep"] = np.packbits(keep, axis=1)
        res[name + "::pmax"] = pr.max(1)
        res[name + "::argmax"] = pr.argmax(1)
        res[name + "::n_keep"] = keep.sum(1)
        print(name, "kept per row", keep.sum(1)[:6])
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
