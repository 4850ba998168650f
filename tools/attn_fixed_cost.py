"""Per-CTA fixed cost of the attention kernels: 2048 queries against Nk keys for growing Nk (time = a + b * Nk when the
fixed part matters). usage: python tools/attn_fixed_cost.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, M, D = 12, 2048, 768
q = torch.randn(B * M, D, device="cuda").bfloat16()
do = torch.randn(B * M, D, device="cuda").bfloat16()
for Nk in (64, 128, 256, 512, 1024, 2048):
    kv = torch.randn(B * Nk, 2 * D, device="cuda").bfloat16()
    dq = torch.empty_like(q); dkv = torch.empty_like(kv)
    meta = ops.attn_ranges(B, M, Nk, device=q.device)
    k, v = kv[:, :D], kv[:, D:]
    for _ in range(2):
        o, lse = ops.attn_fwd(q, k, v, B, H, M, Nk, meta=meta)
        ops.attn_bwd(q, k, v, o, do, lse, B, H, M, Nk, dq, dkv[:, :D], dkv[:, D:], meta=meta)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 10
    e[0].record()
    for _ in range(n): o, lse = ops.attn_fwd(q, k, v, B, H, M, Nk, meta=meta)
    e[1].record()
    for _ in range(n): ops.attn_bwd(q, k, v, o, do, lse, B, H, M, Nk, dq, dkv[:, :D], dkv[:, D:], meta=meta)
    e[2].record()
    torch.cuda.synchronize()
    print("Nk %5d  fwd %.3f ms  bwd %.3f ms (incl. prep + dq cast)" % (Nk, e[0].elapsed_time(e[1]) / n, e[1].elapsed_time(e[2]) / n), flush=True)
