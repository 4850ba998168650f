"""2+ GPUs, launched by torch.distributed.run: one optimisation step of a small 4-modality model under
DistributedDataParallel (each rank its own shard, NCCL all-reduce overlapped with backward) must leave the same weights as
one process stepping on the concatenated global batch ("after-step weights across 1 vs N GPUs on the
same global batch within bf16 tolerance"). Equal per-modality target counts per rank, so the rank-local 'mod' mean averaged
by DDP equals the global mean (egom2p_model.py:639-642)."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist
import synth
from test_model_gpu import build_model


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    # `--one-gpu`: both ranks on cuda:0 with the gloo backend (gradient buckets are reduced through the host, so no kernel of
    # one rank ever waits for a kernel of the other): the same DDP code path -- reducer hooks on the per-block autograd
    # nodes, bucketed all-reduce overlapped with backward -- on a box with a single GPU.
    one_gpu = "--one-gpu" in sys.argv
    local = 0 if one_gpu else local
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if one_gpu:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.make_cfg(384, 6, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=1024, video_thw=(5, 8, 8))
    sd = synth.make_state_dict(cfg, 3)
    per = 2
    kw = dict(n_in={"tok_cam": [15] * per, "tok_depth": [145] * per, "tok_gaze": [15] * per, "tok_rgb": [145] * per},
              n_tgt={"tok_cam": [15] * per, "tok_depth": [145] * per, "tok_gaze": [15] * per, "tok_rgb": [145] * per})
    shards = [synth.make_batch(cfg, B=per, seed=100 + r, **kw) for r in range(world)]
    cu = lambda md: {m: {k: v.to(dev) for k, v in d.items()} for m, d in md.items()}

    grads = {}

    def one_step(net, model, md, tag):
        from egom2p_b200.optim import FusedAdamW
        opt = FusedAdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
        random.seed(7)
        loss, _ = net(cu(md), 320, 320)
        loss.backward()
        grads[tag] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
        opt.clip_grad_norm_(1.0)
        opt.step()
        return float(loss.detach())

    model = build_model(cfg).to(dev)
    model.load_state_dict(sd, strict=True)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False, broadcast_buffers=False)
    loss_ddp = one_step(net, model, shards[rank], "ddp")
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, torch.tensor([loss_ddp], device=dev))
    ok = True
    if rank == 0:
        ref = build_model(cfg).to(dev)
        ref.load_state_dict(sd, strict=True)
        glob = {m: {k: torch.cat([s[m][k] for s in shards], 0) for k in shards[0][m]} for m in shards[0]}
        loss_ref = one_step(ref, ref, glob, "single")
        mean_ddp = float(torch.cat(losses).mean())
        # the all-reduced (averaged) gradients equal the gradients of the single-process step on the global batch; the first
        # Adam step is sign-like (+-lr per element), so weights are compared by how many elements moved the other way
        worst_g, flipped = 0.0, 0.0
        for (n, p), (_, q) in zip(model.named_parameters(), ref.named_parameters()):
            g1, g2 = grads["ddp"][n].float(), grads["single"][n].float()
            worst_g = max(worst_g, ((g1 - g2).norm() / (g2.norm() + 1e-12)).item())
            flipped = max(flipped, ((p.detach() - q.detach()).abs() > 1e-3).float().mean().item())
        print(f"ddp loss (mean over ranks) {mean_ddp:.6f} vs single-process global batch {loss_ref:.6f}; worst relative gradient "
              f"difference {worst_g:.3e}; largest fraction of a tensor's elements stepping the other way {flipped:.3e}", flush=True)
        ok = abs(mean_ddp - loss_ref) < 1e-3 * abs(loss_ref) and worst_g < 2e-2 and flipped < 2e-2
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
