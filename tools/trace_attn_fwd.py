"""Debug: life of one late CTA of the attention forward in cycles. Needs `EGOM2P_TRACE=1 python -m egom2p_b200.build --force`
(the debug build adds clock stamps and the egom2p_debug_attn_fwd_trace export; never ship it)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops, _lib
B, H, M, D = 16, 12, 2048, 768
Nk = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
q = torch.randn(B * M, D, device="cuda").bfloat16()
kv = torch.randn(B * Nk, 2 * D, device="cuda").bfloat16()
meta = ops.attn_ranges(B, M, Nk, device=q.device)
for _ in range(3):
    ops.attn_fwd(q, kv[:, :D], kv[:, D:], B, H, M, Nk, meta=meta)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 32)()
lib = ctypes.CDLL(_lib.LIB_PATH)
assert lib.egom2p_debug_attn_fwd_trace(buf) == 0
names = ["CTA start", "set-up barrier passed", "Q landed", "Q in TMEM / path decided", "S(0) ready", "S(1) ready", "S(last) ready",
         "last P published", "last PV retired", "outputs stored", "CTA end"]
t0 = buf[0]
prev = t0
for i, n in enumerate(names):
    print("%-26s +%7d cycles (step %6d)" % (n, buf[i] - t0, buf[i] - prev))
    prev = buf[i]
