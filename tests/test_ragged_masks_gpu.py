"""GPU: the reference-distribution regime: masks drawn by the UNMODIFIED reference UnifiedMasking
(tests/golden/ref_masks_egob.npz, oracle/gen_golden_masks.py) at ego-b sizes -- budgets 2048 / 2048 over 10300 positions,
very ragged (31 .. 2048 valid inputs). The index plan (kept slots, pads, modality ids, targets, decoder key ranges) must be
bit-exact against the oracle's restatement of forward_mask_encoder / forward_mask_decoder on every sample."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synth  # noqa: E402
from test_kernels_gpu import _plan_case  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from egom2p_b200 import ops as _ops
    return _ops


def test_reference_masks_index_plan_bit_exact(ops):
    import bench
    cfg = synth.make_cfg(12, 1, 0, 0, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=64000)
    g = bench.load_ref_masks()
    for off in (0, 8, 16):
        md = bench.make_batch_ref_masks(8, off, 5 + off, pin=False, g=g)
        md = {m: {k: (v.reshape(v.shape[0], -1) if k == "tensor" else v) for k, v in d.items()} for m, d in md.items()}
        _plan_case(ops, cfg, md, 2048, 2048, ["tok_rgb", "tok_gaze", "tok_depth", "tok_cam"])


def test_reference_masks_flop_count_matches_dense_formula():
    import bench
    d = bench.make_batch(2, 1, pin=False)
    assert abs(bench.step_flops(d) / 2 - bench.FLOP_PER_SAMPLE_STEP) < 1e-3 * bench.FLOP_PER_SAMPLE_STEP
