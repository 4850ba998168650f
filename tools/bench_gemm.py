"""Per-shape timing of the tcgen05 GEMM against cuBLAS (torch.matmul) for the GEMM shapes of one ego-b step."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from egom2p_b200 import ops

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main(b=16):
    R = b * 2048
    D, F = 768, 2048
    shapes = []
    for name, N, K in [("qkv", 3*D, D), ("proj", D, D), ("kv", 2*D, D), ("fc13", 2*F, D), ("fc2", D, F)]:
        shapes.append((name + ".fwd", "tn", R, N, K))
        shapes.append((name + ".dgrad", "nn", R, K, N))
        shapes.append((name + ".wgrad", "tt", N, K, R))
    Rm = b * 1009
    shapes += [("head.dy", "nn", Rm, D, 8192), ("head.dw", "tt", 8192, D, Rm)]
    out = []
    for name, kind, M, N, K in shapes:
        if kind == "tn":
            A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
            C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
            f = lambda: ops.gemm(A, B, M, N, K, out_bf16=C); g = lambda: torch.matmul(A, B.t())
        elif kind == "nn":
            A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16()
            C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
            f = lambda: ops.gemm(A, B, M, N, K, b_mn=True, out_bf16=C); g = lambda: torch.matmul(A, B)
        else:
            A = torch.randn(K, M, device="cuda").bfloat16(); B = torch.randn(K, N, device="cuda").bfloat16()
            C = torch.empty(M, N, dtype=torch.float32, device="cuda")
            f = lambda: ops.gemm(A, B, M, N, K, a_mn=True, b_mn=True, out_f32=C); g = lambda: torch.matmul(A.t(), B)
        t1, t2 = timeit(f), timeit(g)
        fl = 2.0 * M * N * K
        out.append((name, M, N, K, round(t1 * 1e3, 1), round(fl / t1 / 1e9, 0), round(t2 * 1e3, 1), round(fl / t2 / 1e9, 0)))
        print("%-12s M=%6d N=%5d K=%6d  ours %8.1f us %6.0f TF/s | cublas %8.1f us %6.0f TF/s" % out[-1], flush=True)

def swiglu(b=16):
    R, D, F = b * 2048, 768, 2048
    x = torch.randn(R, D, device="cuda").bfloat16(); w13 = torch.randn(2 * F, D, device="cuda").bfloat16() * 0.03
    w2 = torch.randn(D, F, device="cuda").bfloat16() * 0.03; dy = torch.randn(R, D, device="cuda").bfloat16()
    ab, g = ops.gemm_swiglu_fwd(x, w13)
    big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # L2 flush between timed launches
    for name, f, fl in (("fc13+swiglu.fwd", lambda: ops.gemm_swiglu_fwd(x, w13), 2.0 * R * 2 * F * D),
                        ("fc2.dgrad+swiglu'", lambda: ops.gemm_swiglu_bwd(dy, w2, ab), 2.0 * R * F * D)):
        ts = []
        for _ in range(5):
            big.zero_(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[2]
        print("%-20s cold-L2 %8.1f us %6.0f TF/s" % (name, t * 1e3, fl / t / 1e9), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "swiglu":
        swiglu(int(sys.argv[1])); sys.exit(0)
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
