"""Per-CTA fixed cost of the attention backward: Mq queries against 2048 keys for growing Mq (the grid = key tiles stays the
same, each CTA walks Mq / 128 query blocks). usage: python tools/attn_fixed_cost_bwd.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, Nk, D = 12, 2048, 768
kv = torch.randn(B * Nk, 2 * D, device="cuda").bfloat16()
k, v = kv[:, :D], kv[:, D:]
dkv = torch.empty_like(kv)
for Mq in (128, 256, 512, 1024, 2048):
    q = torch.randn(B * Mq, D, device="cuda").bfloat16()
    do = torch.randn(B * Mq, D, device="cuda").bfloat16()
    dq = torch.empty_like(q)
    meta = ops.attn_ranges(B, Mq, Nk, device=q.device)
    o, lse = ops.attn_fwd(q, k, v, B, H, Mq, Nk, meta=meta)
    for _ in range(2):
        ops.attn_bwd(q, k, v, o, do, lse, B, H, Mq, Nk, dq, dkv[:, :D], dkv[:, D:], meta=meta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): ops.attn_bwd(q, k, v, o, do, lse, B, H, Mq, Nk, dq, dkv[:, :D], dkv[:, D:], meta=meta)
    e1.record()
    torch.cuda.synchronize()
    print("Mq %5d (%2d blocks per CTA)  bwd %.3f ms (incl. prep + dq cast)" % (Mq, Mq // 128, e0.elapsed_time(e1) / n), flush=True)
