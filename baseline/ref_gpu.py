"""The bar to beat (SURVEY.md section 8(d) "Reference GPU bar"): the UNMODIFIED reference module from baseline/_ref
(tools/install_reference.py), eager PyTorch under torch.autocast(bf16) exactly as run_training_egom2p.py:725-746 runs it
(forward, loss.backward(), clip_grad_norm_(1.0), AdamW), `egom2p_base_12e_12d_swiglu_nobias` with random-init weights, on
the same dense synthetic batches bench.py uses, timed with CUDA events on the same GPU.

None of this repo's kernels, modules or engine are on this path: it imports only torch and baseline/_ref/egom2p (with
empty stubs for the four packages the reference imports at module level but never calls here: boto3, webdataset,
albumentations, braceexpand -- SURVEY.md section 8(c)).

  python baseline/ref_gpu.py [--batch 4] [--steps 20] [--warmup 5]      -> one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types
from unittest.mock import MagicMock

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
MODS = ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"]   # sorted domains, run_training_egom2p.py:274-276


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return MagicMock(name=f"{self.__name__}.{name}")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "egom2p"))


def import_reference():
    for name in ["boto3", "boto3.s3", "boto3.s3.transfer", "webdataset", "webdataset.handlers", "webdataset.filters",
                 "albumentations", "braceexpand"]:
        if name not in sys.modules:
            m = _Stub(name)
            m.__path__ = []
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import egom2p.models.egom2p_model  # noqa: F401
    from egom2p.data.modality_info import MODALITY_INFO
    from egom2p.utils.timm.model_builder import create_model
    return MODALITY_INFO, create_model


def time_reference(batch: int, steps: int, warmup: int, device: torch.device, make_batch, ddp: bool = False) -> dict:
    """ms / step of the reference training step at `batch` samples on `device`. make_batch(b, seed, pin) -> CPU mod_dict.
    ddp: wrap the model in DistributedDataParallel exactly as run_training_egom2p.py:514 does (process group already up)."""
    MI, create_model = import_reference()
    torch.manual_seed(0)
    model = create_model("egom2p_base_12e_12d_swiglu_nobias",
                         encoder_embeddings={m: MI[m]["encoder_embedding"]() for m in MODS},
                         decoder_embeddings={m: MI[m]["decoder_embedding"]() for m in MODS},
                         modality_info={m: MI[m] for m in MODS}, num_register_tokens=0).to(device)
    model.train()
    net = model
    if ddp:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], find_unused_parameters=False)
    decay = [p for n, p in model.named_parameters() if not ("norm" in n or n.endswith(".bias"))]
    no_decay = [p for n, p in model.named_parameters() if ("norm" in n or n.endswith(".bias"))]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": no_decay, "weight_decay": 0.0}],
                            lr=1e-4, betas=(0.9, 0.95), eps=1e-8, fused=True)
    params = list(model.parameters())
    batches = [{m: {k: v.to(device) for k, v in d.items()} for m, d in make_batch(batch, 1234 + s, False).items()}
               for s in range(4)]

    def step(i):
        md = {m: dict(d) for m, d in batches[i % len(batches)].items()}   # the adapters add keys to the inner dicts
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _ = net(md, 2048, 2048, loss_type="mod")
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / steps
    out = {"batch_per_gpu": batch, "ms_per_step": ms, "tokens_per_s": batch * 4096 / (ms / 1e3), "steps": steps, "warmup": warmup,
           "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30,
           "what": "unmodified reference EgoM2P (baseline/_ref), eager torch.autocast(bf16): fwd + bwd + clip_grad_norm_(1.0) + "
                   "fused AdamW, dense synthetic regime, same GPU"}
    del model, opt, params, batches
    torch.cuda.empty_cache()
    return out


GEN_WORKLOADS = {  # eval_model_rgb2depth.py:45-59, eval_model_rgb2cam.py:40-54, eval_model_rgb2gaze.py:41-55, eval_model_depth2rgb.py:34-48
    "rgb2depth": dict(cond="tok_rgb", target="tok_depth", ntoks=5120, steps=3),
    "rgb2cam": dict(cond="tok_rgb", target="tok_cam", ntoks=30, steps=3),
    "rgb2gaze": dict(cond="tok_rgb", target="tok_gaze", ntoks=30, steps=5),
    "depth2rgb": dict(cond="tok_depth", target="tok_rgb", ntoks=5120, steps=6),
}


def time_reference_generation(workload: str, batch: int, reps: int, warmup: int, device: torch.device) -> dict:
    """Latency of the reference's own generation (GenerationSampler.generate, guided ROAR, T = 0.01, top-p 0.8, CFG 2.0) for one
    batch of `batch` clips, exactly as the eval scripts set it up: fp32 with TF32 matmuls allowed, no autocast, no grad."""
    MI, create_model = import_reference()
    from egom2p.models.generate import (GenerationSampler, build_chained_generation_schedules, init_empty_target_modality,
                                        init_full_input_modality)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    w = GEN_WORKLOADS[workload]
    torch.manual_seed(0)
    model = create_model("egom2p_base_12e_12d_swiglu_nobias",
                         encoder_embeddings={m: MI[m]["encoder_embedding"]() for m in MODS},
                         decoder_embeddings={m: MI[m]["decoder_embedding"]() for m in MODS},
                         modality_info={m: MI[m] for m in MODS}, num_register_tokens=0).eval().to(device)
    sampler = GenerationSampler(model)
    schedule = build_chained_generation_schedules(
        cond_domains=[w["cond"]], target_domains=[w["target"]], tokens_per_target=[w["ntoks"]], autoregression_schemes=["roar"],
        decoding_steps=[w["steps"]], token_decoding_schedules=["linear"], temps=[0.01], temp_schedules=["constant"],
        cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True)
    g = torch.Generator().manual_seed(1)

    def one():
        md = {w["cond"]: {"tensor": torch.randint(0, 64000, (batch, 5, 32, 32), generator=g).to(device),
                          "input_mask": torch.zeros(batch, 5120, dtype=torch.bool, device=device),
                          "target_mask": torch.ones(batch, 5120, dtype=torch.bool, device=device)}}
        md = init_empty_target_modality(md, MI, w["target"], batch, w["ntoks"], device)
        md = init_full_input_modality(md, MI, w["cond"], device)
        out = sampler.generate(md, schedule, verbose=False, seed=0, top_p=0.8, top_k=0.0)
        return out[w["target"]]["tensor"].sum().item()

    with torch.no_grad():
        for _ in range(warmup):
            one()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            one()
        e1.record()
        torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    out = {"workload": workload, "batch": batch, "ms_per_call": ms, "clips_per_s": batch / (ms / 1e3), "reps": reps,
           "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30,
           "what": "unmodified reference GenerationSampler.generate (baseline/_ref), fp32 + TF32, same GPU"}
    del model, sampler
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[4])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:   # torchrun: the reference under DDP, one rank per GPU (the reference's own launch mode)
        import torch.distributed as dist
        rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        sys.path.insert(0, ROOT)
        from bench import make_batch
        res = []
        for b in args.batch:
            r = time_reference(b, args.steps, args.warmup, dev, lambda bb, seed, pin: make_batch(bb, seed + 1000 * rank, pin), ddp=True)
            ms = torch.tensor([r["ms_per_step"]], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            r["ms_per_step"] = float(ms)
            r["n_gpus"], r["tokens_per_s"] = world, world * b * 4096 / (float(ms) / 1e3)
            res.append(r)
        if rank == 0:
            print(json.dumps({"reference_gpu_ddp": res}), flush=True)
        dist.barrier()
        dist.destroy_process_group()
        return
    if not available():
        print(json.dumps({"reference_gpu": None, "unavailable": "baseline/_ref/egom2p missing (run tools/install_reference.py)"}))
        return
    sys.path.insert(0, ROOT)
    from bench import make_batch
    dev = torch.device("cuda", 0)
    res = []
    for b in args.batch:
        try:
            res.append(time_reference(b, args.steps, args.warmup, dev, make_batch))
        except torch.cuda.OutOfMemoryError:
            res.append({"batch_per_gpu": b, "oom": True})
            torch.cuda.empty_cache()
    print(json.dumps({"reference_gpu": res}), flush=True)


if __name__ == "__main__":
    main()
