"""Cold-L2 timing of every GEMM call shape of one ego-b step (b per GPU as argv[1]) through the same ops the model uses,
against torch (cuBLAS) doing the same math incl. the fp32 residual add. L2 is flushed before every timed launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops
b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
R, D, F = b * 2048, 768, 2048
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
bf, f32 = torch.bfloat16, torch.float32


def cold(f, n=5):
    ts = []
    for _ in range(n):
        big.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2] * 1e3


rows = []
def case(name, count, fl, ours, ref):
    t1, t2 = cold(ours), cold(ref)
    rows.append((name, count, t1, t2))
    print("%-22s x%3d  ours %7.1f us %5.0f TF/s | torch %7.1f us %5.0f TF/s" % (name, count, t1, fl / t1 / 1e6, t2, fl / t2 / 1e6), flush=True)


x = torch.randn(R, D, device="cuda").to(bf); res = torch.randn(R, D, device="cuda")
dyb = torch.randn(R, D, device="cuda").to(bf)
for name, N, cnt in (("qkv", 3 * D, 24), ("kv", 2 * D, 12), ("q", D, 12)):
    w = (torch.randn(N, D, device="cuda") * 0.03).to(bf)
    dout = torch.randn(R, N, device="cuda").to(bf)
    case(name + ".fwd", cnt, 2.0 * R * N * D, lambda: ops.linear_fwd(x, w), lambda: torch.matmul(x, w.t()))
    case(name + ".dgrad", cnt, 2.0 * R * N * D, lambda: ops.linear_dgrad(dout, w), lambda: torch.matmul(dout, w))
    case(name + ".wgrad", cnt, 2.0 * R * N * D, lambda: ops.linear_wgrad(dout, x), lambda: torch.matmul(dout.t(), x).float())
wp = (torch.randn(D, D, device="cuda") * 0.03).to(bf)
case("proj.fwd(+res,f32)", 36, 2.0 * R * D * D, lambda: ops.linear_fwd(x, wp, addend=res, out_dtype=f32), lambda: torch.addmm(res, x.float(), wp.float().t()) if False else (torch.matmul(x, wp.t()).float() + res))
case("proj.dgrad", 36, 2.0 * R * D * D, lambda: ops.linear_dgrad(dyb, wp), lambda: torch.matmul(dyb, wp))
case("proj.wgrad", 36, 2.0 * R * D * D, lambda: ops.linear_wgrad(dyb, x), lambda: torch.matmul(dyb.t(), x).float())
g = torch.randn(R, F, device="cuda").to(bf); w2 = (torch.randn(D, F, device="cuda") * 0.03).to(bf)
case("fc2.fwd(+res,f32)", 24, 2.0 * R * D * F, lambda: ops.linear_fwd(g, w2, addend=res, out_dtype=f32), lambda: torch.matmul(g, w2.t()).float() + res)
case("fc2.wgrad", 24, 2.0 * R * D * F, lambda: ops.linear_wgrad(dyb, g), lambda: torch.matmul(dyb.t(), g).float())
w13 = (torch.randn(2 * F, D, device="cuda") * 0.03).to(bf)
ab, gg = ops.gemm_swiglu_fwd(x, w13)
dab = torch.randn(R, 2 * F, device="cuda").to(bf)
case("fc13+swiglu.fwd", 24, 2.0 * R * 2 * F * D, lambda: ops.gemm_swiglu_fwd(x, w13), lambda: torch.matmul(x, w13.t()))
case("fc2.dgrad+swiglu'", 24, 2.0 * R * F * D, lambda: ops.gemm_swiglu_bwd(dyb, w2, ab), lambda: torch.matmul(dyb, w2))
case("fc13.dgrad", 24, 2.0 * R * 2 * F * D, lambda: ops.linear_dgrad(dab, w13), lambda: torch.matmul(dab, w13))
case("fc13.wgrad", 24, 2.0 * R * 2 * F * D, lambda: ops.linear_wgrad(dab, x), lambda: torch.matmul(dab.t(), x).float())
print("step total: ours %.1f ms, torch %.1f ms" % (sum(c * t for _, c, t, _ in rows) / 1e3, sum(c * t for _, c, _, t in rows) / 1e3))
