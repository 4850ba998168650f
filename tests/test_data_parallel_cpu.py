"""CPU, world_size 2, gloo: the host side of the only partition the path has -- data parallel over samples (SURVEY.md
section 8e; run_training_egom2p.py:397-399,514). Checks (1) rank-seeded synthetic shards (bench.make_batch) are
deterministic per rank and differ between ranks, (2) the module is DistributedDataParallel-compatible exactly as the
reference wraps it (find_unused_parameters=False): every parameter -- tied heads and shared modality embeddings counted
once -- goes through the reducer and receives the mean of the per-rank gradients, (3) buffers (fixed sin-cos tables) are
left alone with broadcast_buffers=False. The kernels need a GPU, so the forward is replaced by a CPU surrogate that touches
every parameter; the NCCL path itself is exercised by `bench.py --gpus N` on the GPU box."""
import hashlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _digest(md):
    h = hashlib.sha256()
    for m in sorted(md):
        for k in sorted(md[m]):
            h.update(np.ascontiguousarray(md[m][k].numpy()).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        import synth
        from test_model_gpu import build_model
        # (1) sharding by rank seed
        d0 = _digest(bench.make_batch(2, 1234 + rank * 1000 + 0, pin=False))
        d0b = _digest(bench.make_batch(2, 1234 + rank * 1000 + 0, pin=False))
        assert d0 == d0b, "rank shard is not deterministic"
        gathered = [None] * world
        dist.all_gather_object(gathered, d0)
        assert len(set(gathered)) == world, "ranks drew the same shard"
        # (2) DDP wiring of a small 4-modality module (same constructor path as ego-b)
        cfg = synth.make_cfg(192, 3, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=512, video_thw=(5, 4, 4))
        torch.manual_seed(100 + rank)  # different init per rank: DDP must broadcast rank 0's parameters
        model = build_model(cfg)
        pos_before = {n: b.clone() for n, b in model.named_buffers()}
        params = list(model.parameters())
        assert len({id(p) for p in params}) == len(params)

        def surrogate(self, scale):  # touches every parameter once; d loss / d p = scale
            return sum((p * scale).sum() for p in self.parameters())

        type(model).forward_backup = type(model).forward
        type(model).forward = surrogate
        try:
            net = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=False, broadcast_buffers=False)
            w0 = [p.detach().clone() for p in params]
            ws = [torch.zeros_like(w0[0]) for _ in range(world)]
            dist.all_gather(ws, w0[0])
            assert all(torch.equal(ws[0], w) for w in ws), "parameters were not broadcast from rank 0"
            net(float(rank + 1)).backward()
            expect = sum(range(1, world + 1)) / world
            for n, p in model.named_parameters():
                assert p.grad is not None, n
                assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), n
            # tied parameters: the vocabulary heads alias the decoder token tables, the modality embeddings are shared
            for m in cfg["mods"]:
                dec = model.decoder_embeddings[m]
                assert dec.to_logits.weight is dec.token_emb.weight
                assert dec.mod_emb is model.encoder_embeddings[m].mod_emb
        finally:
            type(model).forward = type(model).forward_backup
        # (3) buffers untouched
        for n, b in model.named_buffers():
            assert torch.equal(b, pos_before[n]), n
        out[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world2_gloo_data_parallel():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.spawn(_worker, args=(world, port, out), nprocs=world, join=False)
    while not ctx.join(timeout=300):  # join() returns after each process exit; True once all have exited cleanly
        pass
    assert dict(out) == {0: "ok", 1: "ok"}
