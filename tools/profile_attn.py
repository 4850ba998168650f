"""Attention fwd + bwd at ego-b encoder shapes for ncu captures. usage: python tools/profile_attn.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, M, D = 12, 2048, 768
qkv = torch.randn(B * M, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * M, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
meta = ops.attn_ranges(B, M, M, device=qkv.device)
for i in range(iters):
    o, lse = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, M, M, meta=meta)
    ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, B, H, M, M, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], meta=meta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
o, lse = ops.attn_fwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, M, M, meta=meta)
e1.record(); torch.cuda.synchronize()
print("fwd ms", e0.elapsed_time(e1), "TF/s", 4.0 * B * H * M * M * 64 / e0.elapsed_time(e1) / 1e9)
e0.record()
ops.attn_bwd(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], o, do, lse, B, H, M, M, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], meta=meta)
e1.record(); torch.cuda.synchronize()
print("bwd ms", e0.elapsed_time(e1), "TF/s", 10.0 * B * H * M * M * 64 / e0.elapsed_time(e1) / 1e9)
