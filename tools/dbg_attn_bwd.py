import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
from egom2p_b200 import ops
from test_attention_bwd_gpu import _ref

def run(B, H, Mq, Nk, mode):
    gen = torch.Generator().manual_seed(Mq * 5 + Nk)
    D = H * 64
    qkv = torch.randn(B, Mq, 3 * D, generator=gen).bfloat16()
    kvsrc = qkv if Mq == Nk else torch.randn(B, Nk, 3 * D, generator=gen).bfloat16()
    do = torch.randn(B, Mq, D, generator=gen).bfloat16()
    q, k, v = qkv[..., :D], kvsrc[..., D:2 * D], kvsrc[..., 2 * D:]
    if mode == "full":
        lo = torch.zeros(B, Mq, dtype=torch.int32); hi = torch.full((B, Mq), Nk, dtype=torch.int32)
    elif mode == "prefix":
        n = torch.tensor([Nk - 3, 0][:B] if B > 1 else [Nk - 3], dtype=torch.int32)
        lo = torch.zeros(B, Mq, dtype=torch.int32); hi = n[:, None].expand(B, Mq).contiguous()
    else:
        bounds = [0, Mq // 2 - 11, Mq - 40, Mq - 25, Mq - 10]
        lo = torch.zeros(B, Mq, dtype=torch.int32); hi = torch.zeros(B, Mq, dtype=torch.int32)
        for a, b_ in zip(bounds[:-1], bounds[1:]):
            lo[:, a:b_] = a; hi[:, a:b_] = b_
    o_ref, dq_ref, dk_ref, dv_ref = _ref(q, k, v, do, lo.long(), hi.long(), H)
    qd, kd = qkv.cuda().reshape(B * Mq, 3 * D), kvsrc.cuda().reshape(B * Nk, 3 * D)
    lod, hid = lo.cuda(), hi.cuda()
    o, lse = ops.attn_fwd(qd[:, :D], kd[:, D:2 * D], kd[:, 2 * D:], B, H, Mq, Nk, lod, hid)
    dqkv = torch.zeros(B * Mq, 3 * D, dtype=torch.bfloat16, device="cuda")
    dq = dqkv[:, :D]
    dkv = torch.zeros(B * Nk, 2 * D, dtype=torch.bfloat16, device="cuda")
    dk, dv = dkv[:, :D], dkv[:, D:]
    ops.attn_bwd(qd[:, :D], kd[:, D:2 * D], kd[:, 2 * D:], o, do.cuda().reshape(B * Mq, D), lse, B, H, Mq, Nk, dq, dk, dv, lod, hid)
    print("case", B, H, Mq, Nk, mode, "lse", lse[:, 0, :4].tolist())
    for name, got, ref in (("dq", dq, dq_ref), ("dk", dk, dk_ref), ("dv", dv, dv_ref)):
        got = got.float().cpu().reshape(ref.shape)
        err = (got - ref.float()).abs()
        scale = ref.abs().max().item()
        bad = (err > 3e-2 * scale + 1e-3)
        print(" ", name, "max err", err.max().item(), "scale", scale, "n_bad", int(bad.sum()))
        if bad.any():
            idx = bad.nonzero()
            print("   bad batches", idx[:, 0].unique().tolist(), "rows min/max", idx[:, 1].min().item(), idx[:, 1].max().item(),
                  "cols min/max", idx[:, 2].min().item(), idx[:, 2].max().item())
            rows = idx[:, 1].unique()
            print("   bad rows", rows[:20].tolist(), "... count", len(rows))
            i = idx[0]
            print("   sample got/ref", got[i[0], i[1], i[2]].item(), ref[i[0], i[1], i[2]].item())

for c in [(2, 2, 300, 517, "prefix"), (1, 1, 300, 517, "full"), (2, 2, 260, 260, "segments"), (1, 1, 20, 24, "prefix"), (1, 2, 512, 512, "full")]:
    run(*c)
