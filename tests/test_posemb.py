"""CPU: positional tables of the package and of the oracle are bit-identical to the reference's
(digests made by oracle/gen_posemb_golden.py from the live reference)."""
import hashlib
import json
import os

import egom2p_oracle as orc
from egom2p_b200 import posemb


def test_posemb_digests(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "posemb_digests.json")))
    for key, ref in g.items():
        parts = key.split("_")
        if parts[0] == "1d":
            n, dim = int(parts[1]), int(parts[2])
            mine = posemb.build_1d_sincos_posemb(n, embed_dim=dim)
            oracle = orc.sincos_1d(n, dim)[None]
        else:
            t, h, w, dim = (int(x) for x in parts[1:])
            mine = posemb.build_3d_sincos_posemb(t, h, w, embed_dim=dim)
            oracle = orc.sincos_3d(t, h, w, dim)[None]
        assert list(mine.shape) == ref["shape"]
        assert hashlib.sha256(mine.numpy().tobytes()).hexdigest() == ref["sha256"], key
        assert hashlib.sha256(oracle.contiguous().numpy().tobytes()).hexdigest() == ref["sha256"], key
