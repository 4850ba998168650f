"""Eager vs whole-step CUDA graph (egom2p_b200.graphed.GraphedTrainStep) at a given per-GPU batch, dense regime.
usage: python tools/bench_graph_step.py [batch] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from egom2p_b200.graphed import GraphedTrainStep
from egom2p_b200.optim import FusedAdamW

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
decay = [p for n, p in model.named_parameters() if not ("norm" in n or n.endswith(".bias"))]
no_decay = [p for n, p in model.named_parameters() if ("norm" in n or n.endswith(".bias"))]
groups = [{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}]
opt = FusedAdamW(groups, lr=1e-4, betas=(0.9, 0.95), eps=1e-8)
batches = [{m: {k: v.to(dev) for k, v in d.items()} for m, d in bench.make_batch(b, 100 + s, False).items()} for s in range(steps + 4)]


def step(md):
    loss, _ = model(md, 2048, 2048)
    loss.backward()
    opt.clip_grad_norm_(1.0)
    opt.step()
    opt.zero_grad(set_to_none=True)


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for i in range(3):
    step(batches[i])
t_eager = timed(lambda i: step(batches[3 + i]), steps)
opt.zero_grad(set_to_none=True)
opt2 = FusedAdamW(groups, lr=1e-4, betas=(0.9, 0.95), eps=1e-8)
runner = GraphedTrainStep(model, opt2, batches[0], 2048, 2048, clip_grad=1.0)
for i in range(3):
    runner(batches[i])
t_graph = timed(lambda i: runner(batches[3 + i]), steps)
print(f"b = {b}: eager {t_eager:.2f} ms / step ({b * 4096 / t_eager:.0f} k tok/s), one CUDA graph {t_graph:.2f} ms / step ({b * 4096 / t_graph:.0f} k tok/s)")
