"""CPU: the masking restatements against the unmodified reference UnifiedMasking (tests/golden/masking_ref.npz, made by
oracle/gen_golden_masking.py): image_mask as a function of its noise (oracle/masking_oracle.py) and the vectorised budget
arithmetic of egom2p_b200.masking.DeviceUnifiedMasking, draw for draw."""
import os

import numpy as np
import torch

import masking_oracle as mo


def test_image_mask_oracle_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "masking_ref.npz"))
    for i in range(int(g["n_image_cases"])):
        L, ib, tb = map(int, g[f"im{i}::cfg"])
        im, tm, attn = mo.image_mask(g[f"im{i}::noise"], ib, tb)
        assert np.array_equal(im, g[f"im{i}::input_mask"]) and np.array_equal(tm, g[f"im{i}::target_mask"]), i
        assert np.array_equal(attn, g[f"im{i}::attn"]), i


def test_budget_arithmetic_matches_reference_draw_for_draw(golden_dir):
    from egom2p_b200.masking import DeviceUnifiedMasking as DM
    g = np.load(os.path.join(golden_dir, "masking_ref.npz"))
    d_in, d_tg = torch.from_numpy(g["draws_in"]), torch.from_numpy(g["draws_tg"])
    n = torch.full((d_in.shape[0],), 2048)
    cap = torch.from_numpy(g["max_tokens"]).int()
    ib = DM._budget_from_draws(d_in[:, 0], d_in[:, 1:], n, cap)
    assert np.array_equal(ib.numpy(), g["budget_in"])
    tb = DM._budget_from_draws(d_tg[:, 0], d_tg[:, 1:], n, torch.maximum(torch.zeros_like(cap), cap - ib))
    assert np.array_equal(tb.numpy(), g["budget_tg"])
