"""One GEMM shape for ncu. usage: python tools/profile_gemm.py M N K [kind=tn|nn|tt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
kind = sys.argv[4] if len(sys.argv) > 4 else "tn"
if kind == "tn":
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
    C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"); f = lambda: ops.gemm(A, B, M, N, K, out_bf16=C)
elif kind == "nn":
    A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16()
    C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda"); f = lambda: ops.gemm(A, B, M, N, K, b_mn=True, out_bf16=C)
else:
    A = torch.randn(K, M, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16()
    C = torch.empty(M, N, dtype=torch.float32, device="cuda"); f = lambda: ops.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out_f32=C)
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f"{kind} M={M} N={N} K={K}: {t*1e3:.1f} us  {2.0*M*N*K/t/1e9:.0f} TF/s")
