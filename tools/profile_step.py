"""One warm-up + N profiled ego-b training steps (dense regime) for ncu launch lists / full captures.
usage: python tools/profile_step.py [batch] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), fused=True)
md = {m: {k: v.to(dev) for k, v in d.items()} for m, d in bench.make_batch(b, 1234, pin=False).items()}
for i in range(1 + steps):
    loss, _ = model(md, 2048, 2048)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step(); opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    print("step", i, float(loss), flush=True)
