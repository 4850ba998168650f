"""Debug: per-iteration timeline of CTA (0,0,0) of the fused attention backward. Needs `EGOM2P_TRACE=1 python -m egom2p_b200.build`."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops, _lib
B, H, M, D = 4, 12, 2048, 768
qkv = torch.randn(B * M, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * M, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
meta = ops.attn_ranges(B, M, M, device=qkv.device)
o, lse = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, M, M, meta=meta)
for _ in range(2):
    ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, B, H, M, M, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], meta=meta)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 64))()
lib = _lib.load()
assert lib.egom2p_debug_attn_trace(buf) == 0
names = ["mmaA:S(j+1)", "mmaB:dV", "mmaA:dP(j+1)", "mmaB:dK", "mmaA:dQ", "math:top", "math:S ready", "math:P done", "math:P stored",
         "math:dP0 loaded", "math:dS stored", "drain:dQ ready", "drain:done", "math:ds0 done", "math:dP1 loaded", "math:ds1 done"]
t0 = buf[5 * 64]
for idx in range(16):
    ev = sorted((buf[s * 64 + idx] - t0, names[s]) for s in range(16) if buf[s * 64 + idx])
    print("it %2d: " % idx + "  ".join("%s@%d" % (n, t) for t, n in ev))

g = lambda i: buf[5 * 64 + i] - buf[5 * 64 + 63]
print("epilogue: dV staged %d | dK staged %d | all warps staged %d" % (g(52), g(51), g(50)))
print("before the barrier: K/V loads issued %d | list built %d | TMEM allocated %d | idle warp at barrier %d | K/V landed %d" % (g(57), g(55), g(54), g(53), g(56)))
print("fixed costs (cycles from CTA start): set-up barrier %d | K/V in TMEM %d | first S^T ready %d | ... | last dS stored %d | dK/dV complete %d | written %d | CTA end %d"
      % (g(62), g(58), buf[6 * 64] - buf[5 * 64 + 63], buf[10 * 64 + 15] - buf[5 * 64 + 63], g(61), g(60), g(59)))
