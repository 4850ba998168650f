"""GPU, T2 of SURVEY.md section 7: the configuration bench.py times -- `egom2p_base_12e_12d_swiglu_nobias` (dim 768,
12 + 12 blocks, 12 heads, 64k / 256 vocabularies), N = M = 2048 -- against outputs of the UNMODIFIED reference (fp32, CPU)
on the same weights, batch, masks and decoder order (tests/golden/egob_{dense,ragged}.npz, oracle/gen_golden_egob.py).

Tolerances are BASELINE.json's: loss within 1e-3 relative, logits within 2e-2 max-abs (all four heads, incl. the
256-vocabulary cam / gaze heads on their full rows), gradients within 3e-2 (5e-2 on the cross-attention query path, see
tests/test_model_gpu.py) -- per-parameter norms for all 245 parameters, relative Frobenius error for the stored ones."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import synth  # noqa: E402
import gen_golden_egob as gg  # noqa: E402  (batch builders only; the reference is not imported at module level)

MODS = gg.MODS


@pytest.fixture(scope="module")
def egob_model():
    import egom2p_b200 as e
    from egom2p_b200.modality_info import MODALITY_INFO as MI
    model = e.create_model("egom2p_base_12e_12d_swiglu_nobias",
                           encoder_embeddings={k: MI[k]["encoder_embedding"]() for k in MODS},
                           decoder_embeddings={k: MI[k]["decoder_embedding"]() for k in MODS},
                           modality_info={k: MI[k] for k in MODS}, num_register_tokens=0)
    model.load_state_dict(synth.make_state_dict(gg.egob_cfg(), gg.SD_SEED), strict=True)
    return model.cuda()


def _to_cuda(md):
    return {m: {k: v.cuda() for k, v in d.items()} for m, d in md.items()}


@pytest.mark.parametrize("case", ["dense", "ragged"])
def test_egob_step_matches_reference(egob_model, golden_dir, case):
    model = egob_model
    g = np.load(os.path.join(golden_dir, f"egob_{case}.npz"))
    cfg = gg.egob_cfg()
    md = gg.dense_batch(cfg) if case == "dense" else gg.ragged_batch(cfg)
    model.zero_grad(set_to_none=True)
    random.seed(gg.SHUFFLE_SEED)
    loss, mod_loss = model(_to_cuda(md), gg.N_ENC, gg.N_DEC, loss_type="mod")
    loss.backward()
    torch.cuda.synchronize()
    ref = float(g["loss"])
    assert abs(loss.item() - ref) / abs(ref) < 1e-3, (loss.item(), ref)
    for m, r in zip(MODS, g["mod_loss"]):
        assert abs(mod_loss[m].item() - r) <= 1e-3 * max(1.0, abs(r)), (m, mod_loss[m].item(), r)
    # ---- gradients: every parameter's norm, and the stored gradients element by element
    norms = dict(zip(g["grad_names"], g["grad_norms"]))
    bad = []
    for n, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        got = p.grad.double().norm().item()
        tol = 5e-2 if ("cross_attn.q." in n or "query_norm" in n) else 3e-2
        if abs(got - norms[n]) > tol * norms[n] + 1e-9:
            bad.append((n, got, norms[n]))
        key = "grad::" + n
        if key in g.files:
            gr = torch.from_numpy(g[key])
            err = ((p.grad.float().cpu() - gr).norm() / (gr.norm() + 1e-12)).item()
            if err > tol:
                bad.append((n, "frobenius", err))
    assert not bad, bad
    # ---- logits (fp32 from TMEM) of every head on its own valid target rows
    model.zero_grad(set_to_none=True)
    with torch.no_grad():
        random.seed(gg.SHUFFLE_SEED)
        logits = model(_to_cuda(md), gg.N_ENC, gg.N_DEC, return_logits=True)
    cols = torch.from_numpy(g["cols"]).cuda()
    worst = {}
    for m in MODS:
        rows = torch.from_numpy(g[f"rows::{m}"].astype(np.int64)).cuda()
        if rows.numel() == 0:
            continue
        lg = logits[m][0].index_select(0, rows).float()
        want = torch.from_numpy(g[f"logits::{m}"]).cuda()
        got = lg if lg.shape[-1] <= 256 else lg.index_select(1, cols)
        worst[m] = (got - want).abs().max().item()
        lse = torch.logsumexp(lg.double(), -1).cpu().numpy()
        worst[m + "/lse"] = float(np.abs(lse - g[f"lse::{m}"]).max())
    assert all(v < 2e-2 for v in worst.values()), worst


def test_egob_index_plan_bit_exact(egob_model, golden_dir):
    """Target ids and modality ids of the compacted decoder sequence == the reference's forward_mask_decoder output."""
    from egom2p_b200 import ops
    g = np.load(os.path.join(golden_dir, "egob_ragged.npz"))
    cfg = gg.egob_cfg()
    md = _to_cuda(gg.ragged_batch(cfg))
    order = [str(x) for x in g["dec_order"]]
    info = egob_model.modality_info
    dp = ops.index_plan([md[m]["target_mask"] for m in order], [info[m]["id"] for m in order], gg.N_DEC, decoder=True,
                        attn_cnt=[md[m]["decoder_attention_mask"].to(torch.int32) for m in order],
                        ids=[md[m]["tensor"].reshape(1, -1) for m in order])
    assert np.array_equal(dp.mod_mask.cpu().numpy(), g["dec_mod"])
    assert np.array_equal(dp.target_ids.cpu().numpy(), g["target_ids"])
