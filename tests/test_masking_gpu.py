"""GPU: device-side masking (egom2p_image_masks + DeviceUnifiedMasking). Bit-exact masks against the reference's image_mask
for the same noise; permutation / budget invariants and uniformity on the Philox path; a masked batch trains."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_image_masks_bit_exact_for_injected_noise(golden_dir):
    from egom2p_b200.masking import image_masks
    g = np.load(os.path.join(golden_dir, "masking_ref.npz"))
    for L in (5120, 30):
        idx = [i for i in range(int(g["n_image_cases"])) if int(g[f"im{i}::cfg"][0]) == L]
        noise = torch.from_numpy(np.stack([g[f"im{i}::noise"] for i in idx])).cuda()
        ib = torch.tensor([int(g[f"im{i}::cfg"][1]) for i in idx], dtype=torch.int32).cuda()
        tb = torch.tensor([int(g[f"im{i}::cfg"][2]) for i in idx], dtype=torch.int32).cuda()
        im, tm, cnt = image_masks(len(idx), L, ib, tb, noise=noise)
        for j, i in enumerate(idx):
            assert np.array_equal(im[j].cpu().numpy(), g[f"im{i}::input_mask"]), i
            assert np.array_equal(tm[j].cpu().numpy(), g[f"im{i}::target_mask"]), i
            assert np.array_equal(cnt[j].cpu().numpy(), g[f"im{i}::attn"]), i


def test_philox_masks_are_uniform_random_subsets():
    from egom2p_b200.masking import image_masks
    B, L = 512, 5120
    ib = torch.randint(0, 2049, (B,), dtype=torch.int32, device="cuda")
    tb = torch.randint(0, 2049, (B,), dtype=torch.int32, device="cuda")
    im, tm, cnt = image_masks(B, L, ib, tb)
    assert torch.equal((~im).sum(1).int(), ib) and torch.equal((~tm).sum(1).int(), tb)          # exact budgets
    assert not bool((~im & ~tm).any())                                                          # inputs and targets are disjoint
    first = torch.where(tb > 0, (tm.float() + torch.arange(L, device="cuda") * 1e-6).argmin(1), 0)
    assert torch.equal(cnt.sum(1), tb) and torch.equal(cnt.gather(1, first[:, None])[:, 0], tb)
    # every position is an input equally often: frequency over the batch ~ mean(ib) / L (binomial tolerance), and two
    # launches draw different masks
    freq = (~im).float().mean(0)
    p = float(ib.float().mean()) / L
    assert float((freq - p).abs().max()) < 6 * (p * (1 - p) / B) ** 0.5
    im2, _, _ = image_masks(B, L, ib, tb)
    assert not torch.equal(im, im2)
    # neighbouring samples are decorrelated: overlap of their input sets ~ product of their densities
    a, b = (~im[0::2]).float(), (~im[1::2]).float()
    assert abs(float((a * b).mean()) - float(a.mean()) * float(b.mean())) < 5e-3


def test_device_masking_feeds_a_training_step(golden_dir):
    import synth
    from egom2p_b200.masking import DeviceUnifiedMasking
    from test_model_gpu import build_model
    cfg = synth.make_cfg(192, 3, 1, 1, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=512, video_thw=(5, 4, 4))
    model = build_model(cfg).cuda()
    model.load_state_dict(synth.make_state_dict(cfg, 2), strict=True)
    alphas = [0.01, 0.1, 1.0, 10.0]
    info = {m: {"type": inf["type"], "max_tokens": inf["len"], "min_tokens": 0, "input_alphas": alphas, "target_alphas": alphas}
            for m, inf in cfg["mods"].items()}
    masking = DeviceUnifiedMasking(info, input_tokens_range=(64, 64), target_tokens_range=(48, 48), sampling_weights=[1, 1, 1, 1])
    B = 16
    ib, tb = masking.token_budgets(B)
    cap = masking.max_tokens
    assert bool((ib <= cap).all()) and bool((tb <= cap - ib).all()) and bool((ib.sum(1) <= 64).all()) and bool((tb.sum(1) <= 48).all())
    tokens = {m: torch.randint(0, inf["vocab"], (B, *(inf["thw"] if "thw" in inf else (inf["len"],))), device="cuda")
              for m, inf in cfg["mods"].items()}
    md = masking(tokens, budgets=(ib, tb))
    for j, m in enumerate(cfg["mods"]):
        assert torch.equal((~md[m]["input_mask"]).sum(1).int(), ib[:, j]) and torch.equal((~md[m]["target_mask"]).sum(1).int(), tb[:, j])
    loss, mod_loss = model(md, 64, 48)
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    # the reference's own budgets for the ego-b mixture (tests/golden/ref_masks_egob.npz, 64 samples): same regime of raggedness
    g = np.load(os.path.join(golden_dir, "ref_masks_egob.npz"))
    info_b = {m: {"type": "img", "max_tokens": L, "min_tokens": 0, "input_alphas": alphas, "target_alphas": alphas}
              for m, L in (("tok_cam", 30), ("tok_depth", 5120), ("tok_gaze", 30), ("tok_rgb", 5120))}
    mb = DeviceUnifiedMasking(info_b, (2048, 2048), (2048, 2048), sampling_weights=[1, 1, 1, 1])
    ib, tb = mb.token_budgets(4096)
    ref_in, ref_tg = g["valid"][:, 0].mean(), g["valid"][:, 1].mean()
    assert abs(float(ib.sum(1).float().mean()) - ref_in) < 0.15 * ref_in + 3 * g["valid"][:, 0].std() / 8
    assert abs(float(tb.sum(1).float().mean()) - ref_tg) < 0.15 * ref_tg + 3 * g["valid"][:, 1].std() / 8
