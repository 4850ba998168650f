"""CPU: pins oracle/egom2p_oracle.py against outputs of the unmodified reference (tests/golden/, made by
oracle/gen_golden.py). Index plan bit-exact; fp32 tensors/loss/grads to round-off (1e-5 rel)."""
import os

import numpy as np
import torch

import egom2p_oracle as orc
import synth

SMALL4_KW = dict(B=4, seed=11,
                 n_in={"tok_cam": [5, 0, 30, 2], "tok_depth": [30, 10, 0, 1], "tok_gaze": [4, 0, 30, 0], "tok_rgb": [25, 40, 4, 0]},
                 n_tgt={"tok_cam": [10, 30, 0, 1], "tok_depth": [20, 0, 40, 0], "tok_gaze": [3, 0, 0, 0], "tok_rgb": [15, 18, 70, 0]})


def small4_cfg():
    return synth.make_cfg(48, 2, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=128, video_thw=(5, 4, 4))


def test_small4_plan_and_embeddings_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "small4.npz"))
    cfg = small4_cfg()
    sd = synth.make_state_dict(cfg, 5)
    md = synth.make_batch(cfg, **SMALL4_KW)
    out = orc.forward(sd, cfg, md, 64, 48, dec_order=list(g["dec_order"]), keep=True)
    assert np.array_equal(out["enc_plan"]["mask"], g["enc_mask"])
    assert np.array_equal(out["enc_plan"]["mod_mask"], g["enc_mod"])
    assert np.array_equal(out["dec_plan"]["mask"], g["dec_mask"])
    assert np.array_equal(out["dec_plan"]["mod_mask"], g["dec_mod"])
    assert np.array_equal(out["dec_plan"]["target_ids"], g["target_ids"])
    assert np.array_equal(np.packbits(out["dec_plan"]["attn_mask"], axis=-1), g["dec_attn"])
    # gathered embeddings: pure fp32 gathers/adds -> bit-exact
    assert np.array_equal(out["enc_x0"].numpy(), g["enc_x0"])
    assert np.array_equal(out["enc_emb"].numpy(), g["enc_emb"])
    assert np.array_equal(out["dec_y0"].numpy(), g["dec_y0"])


def test_small4_loss_logits_grads(golden_dir):
    g = np.load(os.path.join(golden_dir, "small4.npz"))
    cfg = small4_cfg()
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("pos_emb") and not (k.endswith(".bias") and "proj_context" not in k))
          for k, v in synth.make_state_dict(cfg, 5).items()}
    for m in cfg["mods"]:  # re-tie after cloning
        sd[f"decoder_embeddings.{m}.to_logits.weight"] = sd[f"decoder_embeddings.{m}.token_emb.weight"]
        sd[f"decoder_embeddings.{m}.mod_emb"] = sd[f"encoder_embeddings.{m}.mod_emb"]
    md = synth.make_batch(cfg, **SMALL4_KW)
    order = list(g["dec_order"])
    for lt in ("mod", "token"):
        out = orc.forward(sd, cfg, md, 64, 48, dec_order=order, loss_type=lt)
        np.testing.assert_allclose(out["loss"].item(), g[f"loss_{lt}"], rtol=2e-6)
        np.testing.assert_allclose([out["mod_loss"][m].item() for m in cfg["mods"]], g[f"mod_loss_{lt}"], rtol=2e-6)
    out = orc.forward(sd, cfg, md, 64, 48, dec_order=order)
    out["loss"].backward()
    norms = dict(zip(g["grad_names"], g["grad_norms"]))
    for n, ref in norms.items():
        if n not in sd or sd[n].grad is None:
            assert ref == 0.0 or n.startswith("decoder_embeddings") and n.endswith("mod_emb"), n
            continue
        np.testing.assert_allclose(sd[n].grad.double().norm().item(), ref, rtol=1e-4, atol=1e-9, err_msg=n)
    for k in g.files:
        if k.startswith("grad::"):
            np.testing.assert_allclose(sd[k[6:]].grad.numpy(), g[k], rtol=1e-4, atol=1e-7, err_msg=k)
    with torch.no_grad():
        lo = orc.forward(sd, cfg, md, 64, 48, dec_order=order, return_logits=True)["logits"]
    for m in cfg["mods"]:
        np.testing.assert_allclose(lo[m][:, :, :16].numpy(), g[f"logits_head::{m}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(torch.logsumexp(lo[m].double(), -1).numpy(), g[f"logits_lse::{m}"], rtol=1e-6)


def test_c1_tiny_fm(golden_dir):
    """BASELINE configs[0]: tiny FM (2 enc / 2 dec layers, dim 256) fwd/bwd on example_data/token cam+gaze tokens."""
    g = np.load(os.path.join(golden_dir, "c1_tiny_fm.npz"))
    cfg = synth.make_cfg(256, 4, 2, 2, ["tok_cam", "tok_gaze"])
    sd = {k: v.clone().requires_grad_(k.endswith("weight") or k.endswith("mod_emb") or k in ("mask_token", "decoder_proj_context.bias"))
          for k, v in synth.make_state_dict(cfg, 7, tie=False).items()}
    for m in cfg["mods"]:
        sd[f"decoder_embeddings.{m}.mod_emb"] = sd[f"encoder_embeddings.{m}.mod_emb"]
    md = synth.make_batch(cfg, B=1, seed=21, n_in={"tok_cam": [14], "tok_gaze": [10]}, n_tgt={"tok_cam": [8], "tok_gaze": [12]})
    md["tok_cam"]["tensor"] = torch.from_numpy(g["cam_ids"])
    md["tok_gaze"]["tensor"] = torch.from_numpy(g["gaze_ids"])
    out = orc.forward(sd, cfg, md, 24, 20, dec_order=list(g["dec_order"]))
    np.testing.assert_allclose(out["loss"].item(), g["loss"], rtol=2e-6)
    np.testing.assert_allclose([out["mod_loss"][m].item() for m in cfg["mods"]], g["mod_loss"], rtol=2e-6)
    out["loss"].backward()
    for n, ref in zip(g["grad_names"], g["grad_norms"]):
        got = 0.0 if sd[n].grad is None else sd[n].grad.double().norm().item()
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-9, err_msg=n)


def test_plan_egob_shapes(golden_dir):
    """ego-b shaped index plan (L = 10300, budgets 2048): stable partition == the reference's fp32 argsort."""
    g = np.load(os.path.join(golden_dir, "plan_egob.npz"))
    cfg = synth.make_cfg(12, 1, 0, 0, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=64000)
    n_in = {"tok_cam": [15, 0, 30, 0], "tok_depth": [1009, 694, 0, 30], "tok_gaze": [15, 0, 30, 30], "tok_rgb": [1009, 0, 5120, 0]}
    n_tg = {"tok_cam": [15, 30, 0, 1], "tok_depth": [1009, 0, 2048, 0], "tok_gaze": [15, 0, 0, 0], "tok_rgb": [1009, 2018, 0, 0]}
    md = synth.make_batch(cfg, B=4, seed=31, n_in=n_in, n_tgt=n_tg)
    mods = list(cfg["mods"])
    ids = {m: cfg["mods"][m]["id"] for m in mods}
    ep = orc.plan_encoder({m: md[m]["input_mask"].numpy() for m in mods}, ids, 2048)
    assert np.array_equal(ep["ids_keep"].astype(np.int32), g["enc_keep"])
    assert np.array_equal(ep["mask"], g["enc_mask"]) and np.array_equal(ep["mod_mask"], g["enc_mod"])
    dp = orc.plan_decoder({m: md[m]["target_mask"].numpy() for m in mods},
                          {m: md[m]["decoder_attention_mask"].numpy() for m in mods},
                          {m: md[m]["tensor"].reshape(4, -1).numpy() for m in mods}, ids, list(g["dec_order"]), 2048)
    assert np.array_equal(dp["mask"], g["dec_mask"]) and np.array_equal(dp["mod_mask"], g["dec_mod"])
    assert np.array_equal(dp["target_ids"], g["target_ids"])
    assert np.array_equal(np.packbits(dp["attn_mask"], axis=-1), g["dec_attn"])


def test_sampling_oracle_matches_reference_filter(golden_dir):
    """oracle/sampling_oracle.py against what the unmodified reference's top_k_top_p_filtering / softmax keep and weigh
    (tests/golden/sampling_filter.npz, generate.py:332-371)."""
    import gen_golden_sampling as ggs
    import sampling_oracle as so
    g = np.load(os.path.join(golden_dir, "sampling_filter.npz"))
    for name, rows, V, scale, top_k, top_p, temp in ggs.CASES:
        lg = ggs.make_logits(name, rows, V, scale)
        keep = so.kept_mask(lg, top_k, top_p)
        want = np.unpackbits(g[name + "::keep"], axis=1)[:, :V].astype(bool)
        # a token whose cumulative mass sits within float32 rounding of top_p may fall on either side
        assert (keep != want).sum(1).max() <= 1, name
        pr = so.probs(lg, temp, top_k, top_p)
        assert np.array_equal(pr.argmax(1), g[name + "::argmax"]), name
        np.testing.assert_allclose(pr.max(1), g[name + "::pmax"], rtol=2e-3)
