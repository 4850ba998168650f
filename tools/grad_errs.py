"""Prints the largest relative gradient errors (B200 module vs CPU oracle) of the small ragged 4-modality step."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import synth
import test_model_gpu as T

cfg = synth.make_cfg(192, 3, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=512, video_thw=(5, 4, 4))
md = synth.make_batch(cfg, B=4, seed=11,
                      n_in={"tok_cam": [5, 0, 30, 2], "tok_depth": [30, 10, 0, 1], "tok_gaze": [4, 0, 30, 0], "tok_rgb": [25, 40, 4, 0]},
                      n_tgt={"tok_cam": [10, 30, 0, 1], "tok_depth": [20, 0, 40, 0], "tok_gaze": [3, 0, 0, 0], "tok_rgb": [15, 18, 70, 0]})
for seed, sseed in [(5, 3), (6, 4), (7, 5)]:
    sd = synth.make_state_dict(cfg, seed)
    model = T.build_model(cfg).cuda()
    model.load_state_dict(sd, strict=True)
    mods = list(cfg["mods"])
    random.seed(sseed)
    order = random.sample(mods, len(mods))
    ref, leaf = T.oracle_run(sd, cfg, md, 64, 48, order)
    random.seed(sseed)
    loss, _ = model(T.to_cuda(md), 64, 48)
    loss.backward()
    errs = []
    for name, p in model.named_parameters():
        g_ref = leaf[name].grad
        if g_ref is None or p.grad is None:
            continue
        g = p.grad.float().cpu()
        errs.append((float((g - g_ref).norm() / (g_ref.norm() + 1e-12)), name, float(g_ref.norm())))
    errs.sort(reverse=True)
    print("seed", seed, "loss", loss.item(), ref["loss"].item())
    for e in errs[:8]:
        print("  %.4f %-45s |g|=%.3e" % e)
