"""Attention fwd + bwd at the ego-b shapes of one step (dense regime) for timing / ncu captures.
usage: python tools/profile_attn.py [B] [iters] [mode]   mode: full (encoder self / decoder cross) | seg (decoder self) | all"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
H, M, D = 12, 2048, 768
qkv = torch.randn(B * M, 3 * D, device="cuda").bfloat16()
do = torch.randn(B * M, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)


def run(name, meta, pairs):
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    for i in range(iters):
        o, lse = ops.attn_fwd(q, k, v, B, H, M, M, meta=meta)
        ops.attn_bwd(q, k, v, o, do, lse, B, H, M, M, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], meta=meta)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    n = 5
    e[0].record()
    for _ in range(n):
        o, lse = ops.attn_fwd(q, k, v, B, H, M, M, meta=meta)
    e[1].record()
    for _ in range(n):
        ops.attn_bwd(q, k, v, o, do, lse, B, H, M, M, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], meta=meta)
    e[2].record()
    torch.cuda.synchronize()
    tf, tb = e[0].elapsed_time(e[1]) / n, e[1].elapsed_time(e[2]) / n
    fl = 4.0 * B * H * pairs * 64
    print("%-5s fwd %.3f ms %6.0f TF/s | bwd %.3f ms %6.0f TF/s" % (name, tf, fl / tf / 1e9, tb, 2.5 * fl / tb / 1e9), flush=True)


if mode in ("full", "all"):
    run("full", ops.attn_ranges(B, M, M, device=qkv.device), M * M)
if mode in ("seg", "all"):  # decoder self-attention of the dense regime: modality segments 1009 / 1009 / 15 / 15
    bounds = [0, 1009, 2018, 2033, 2048]
    lo = torch.zeros(B, M, dtype=torch.int32)
    hi = torch.zeros(B, M, dtype=torch.int32)
    for a, b_ in zip(bounds[:-1], bounds[1:]):
        lo[:, a:b_] = a
        hi[:, a:b_] = b_
    run("seg", ops.attn_ranges(B, M, M, lo.cuda(), hi.cuda()), sum((b_ - a) ** 2 for a, b_ in zip(bounds[:-1], bounds[1:])))
