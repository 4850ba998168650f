"""TEST INFRASTRUCTURE ONLY. SHA-256 digests + sampled rows of the reference's own positional tables
(authoring container only) -> tests/golden/posemb_digests.json."""
import hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference
import_reference()
from egom2p.models.egom2p_utils import build_1d_sincos_posemb, build_3d_sincos_posemb
out = {}
for dim in (48, 192, 256, 384, 768):
    t = build_1d_sincos_posemb(30, embed_dim=dim)
    out[f"1d_30_{dim}"] = {"sha256": hashlib.sha256(t.numpy().tobytes()).hexdigest(), "shape": list(t.shape)}
for (tt, h, w, dim) in ((5, 32, 32, 768), (5, 32, 32, 384), (5, 4, 4, 48), (5, 8, 8, 192)):
    t = build_3d_sincos_posemb(tt, h, w, embed_dim=dim)
    out[f"3d_{tt}_{h}_{w}_{dim}"] = {"sha256": hashlib.sha256(t.numpy().tobytes()).hexdigest(), "shape": list(t.shape)}
json.dump(out, open(os.path.join(os.path.dirname(HERE), "tests", "golden", "posemb_digests.json"), "w"), indent=1)
print(out)
