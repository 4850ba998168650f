"""How a persistent GEMM behaves when another resident kernel holds a few SMs (the situation under DDP, where NCCL's
all-reduce kernels run beside the backward GEMMs): one spinning CTA on a side stream vs. none.
usage: python tools/bench_gemm_corun.py [M N K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 2304, 768)
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
f = lambda: ops.gemm(A, B, M, N, K, out_bf16=C)
g = lambda: torch.matmul(A, B.t(), out=C)
side = torch.cuda.Stream()


def timed(fn, spin_blocks):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    for _ in range(spin_blocks):
        with torch.cuda.stream(side):
            torch.cuda._sleep(40_000_000)   # ~20 ms, one CTA each call (serialised on the side stream: still one SM at a time)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


for name, fn in (("ours", f), ("cublas", g)):
    t0 = timed(fn, 0)
    t1 = timed(fn, 1)
    print(f"{name}: alone {t0:.1f} us | with one SM held by another kernel {t1:.1f} us  (x{t1 / t0:.2f})")
