"""LayerNorm fwd / bwd timing at the ego-b step shape (rows = b * 2048, dim 768), cold L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
R, D = b * 2048, 768
x = torch.randn(R, D, device="cuda"); w = torch.ones(D, device="cuda")
dy = torch.randn(R, D, device="cuda").bfloat16(); dxin = torch.randn(R, D, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
h, _, mean, rstd = ops.layernorm_fwd(x, w, 1e-6)
dw = torch.zeros(D, device="cuda")
def cold(f, n=7):
    ts = []
    for _ in range(n):
        big.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2] * 1e3
t = cold(lambda: ops.layernorm_fwd(x, w, 1e-6)); print("ln_fwd  %.1f us  %.0f GB/s" % (t, R * D * 6 / t / 1e3))
t = cold(lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dx_in=dxin, d_weight=dw, want_bf16=True)); print("ln_bwd  %.1f us  %.0f GB/s (x, dy, dx_in in; dx, bf16 out)" % (t, R * D * 16 / t / 1e3))
