"""CPU: the drop-in boundary -- state_dict / parameter layout identical to the reference's ego-b (manifest made from the
live reference), C-ABI library loads and exports every symbol declared in include/egom2p_b200.h, host-side API
behaviour (errors, freeze helpers, registry). No kernel is launched here."""
import ctypes
import json
import os
import re

import pytest
import torch

import egom2p_b200 as e
from egom2p_b200 import _lib
from egom2p_b200.modality_info import MODALITY_INFO as MI

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(name="egom2p_base_12e_12d_swiglu_nobias", mods=None, **kw):
    mods = mods or e.MOD4
    return e.create_model(name, encoder_embeddings={k: MI[k]["encoder_embedding"]() for k in mods},
                          decoder_embeddings={k: MI[k]["decoder_embedding"]() for k in mods},
                          modality_info={k: MI[k] for k in mods}, num_register_tokens=0, **kw)


@pytest.fixture(scope="module")
def egob():
    torch.manual_seed(0)
    return build()


def test_state_dict_matches_reference_manifest(egob, golden_dir):
    man = json.load(open(os.path.join(golden_dir, "egob_state_dict_manifest.json")))
    sd = egob.state_dict()
    assert list(sd.keys()) == list(man["state_dict"].keys())          # same keys, same order
    assert all(list(sd[k].shape) == man["state_dict"][k] for k in sd)
    assert [n for n, _ in egob.named_parameters()] == man["named_parameters"]
    assert sum(p.numel() for p in egob.parameters()) == man["n_params"] == 396226560


def test_tying_and_buffers(egob):
    for m in e.MOD4:
        dec, enc = egob.decoder_embeddings[m], egob.encoder_embeddings[m]
        assert dec.to_logits.weight is dec.token_emb.weight        # tied head (decoder_embeddings.py:447-449)
        assert dec.mod_emb is enc.mod_emb                          # shared modality embedding (egom2p_model.py:179-183)
    bufs = dict(egob.named_buffers())
    assert len(bufs) == 82 and sum(k.endswith("pos_emb") for k in bufs) == 8
    assert egob.no_weight_decay() == set()
    assert egob.decoder_proj_context.bias is not None and float(egob.decoder_proj_context.bias.abs().sum()) == 0.0


def test_modality_ids():
    assert {k: v["id"] for k, v in MI.items()} == {"tok_rgb": 7613, "tok_depth": 6323, "tok_cam": 349, "tok_gaze": 26680}


def test_freeze_helpers(egob):
    egob.freeze_shared_params()
    assert not any(p.requires_grad for p in egob.encoder.parameters())
    assert all(p.requires_grad for p in egob.encoder_embeddings.parameters())
    egob.freeze_params_except_specific_embeddings("tok_rgb-tok_cam")
    assert not egob.encoder_embeddings["tok_rgb"].token_emb.weight.requires_grad
    assert egob.encoder_embeddings["tok_depth"].token_emb.weight.requires_grad
    egob.unfreeze_all()
    assert all(p.requires_grad for p in egob.parameters())


def test_fails_loudly_without_gpu_and_on_bad_args(egob):
    md = {m: {"tensor": torch.zeros(1, 30, dtype=torch.int64), "input_mask": torch.zeros(1, 30, dtype=torch.bool),
              "target_mask": torch.ones(1, 30, dtype=torch.bool), "decoder_attention_mask": torch.zeros(1, 30, dtype=torch.int32)}
          for m in ("tok_cam", "tok_gaze")}
    with pytest.raises(ValueError):
        egob(md, 8, 8, loss_type="nope")
    with pytest.raises(RuntimeError, match="CUDA"):
        egob(md, 8, 8)            # CPU tensors: there is no fallback path
    with pytest.raises(NotImplementedError):
        build(qk_norm=True)


def test_registry_names():
    from egom2p_b200 import registry
    for n in ("egom2p_tiny_6e_6d_swiglu_nobias", "egom2p_small_8e_8d_swiglu_nobias", "egom2p_base_12e_12d_swiglu_nobias",
              "egom2p_base_12e_12d_swiglu_nobias_causal"):
        assert registry.is_model(n)
    with pytest.raises(RuntimeError):
        e.create_model("egom2p_nonexistent")


def test_c_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "egom2p_b200.h")).read()
    declared = set(re.findall(r"\b(egom2p_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _lib.load().egom2p_abi_version() == 1
    assert _lib.load().egom2p_attn_lse_stride(2048) == 2048 and _lib.load().egom2p_attn_lse_stride(20) == 128


@pytest.mark.skipif(not os.path.isdir("/root/reference/egom2p"), reason="the reference tree exists in the authoring container only")
def test_register_into_reference_overrides_registry():
    """INTEGRATION.md's hook: after register_into_reference() the reference's own create_model (the call get_model makes,
    run_training_egom2p.py:381-387) builds the B200 module under the unchanged names, with REFERENCE adapter objects, and a
    reference state_dict loads strictly into it."""
    from _ref_import import import_reference
    import_reference()
    from egom2p.data.modality_info import MODALITY_INFO as REF_MI
    from egom2p.utils.timm.model_builder import create_model as ref_create
    e.register_into_reference()
    mods = ["tok_cam", "tok_gaze"]
    model = ref_create("egom2p_tiny_6e_6d_swiglu_nobias",
                       encoder_embeddings={m: REF_MI[m]["encoder_embedding"]() for m in mods},
                       decoder_embeddings={m: REF_MI[m]["decoder_embedding"]() for m in mods},
                       modality_info={m: REF_MI[m] for m in mods}, num_register_tokens=0)
    assert type(model).__module__ == "egom2p_b200.model" and len(model.encoder) == 6
    assert type(model.encoder_embeddings["tok_cam"]).__module__.startswith("egom2p.models")   # reference adapters, duck-typed
    ours = e.create_model("egom2p_tiny_6e_6d_swiglu_nobias",
                          encoder_embeddings={m: MI[m]["encoder_embedding"]() for m in mods},
                          decoder_embeddings={m: MI[m]["decoder_embedding"]() for m in mods},
                          modality_info={m: MI[m] for m in mods}, num_register_tokens=0)
    assert list(ours.state_dict().keys()) == list(model.state_dict().keys())
    model.load_state_dict(ours.state_dict(), strict=True)
    causal = ref_create("egom2p_base_12e_12d_swiglu_nobias_causal",
                        encoder_embeddings={m: REF_MI[m]["encoder_embedding"]() for m in mods},
                        decoder_embeddings={m: REF_MI[m]["decoder_embedding"]() for m in mods},
                        modality_info={m: REF_MI[m] for m in mods}, num_register_tokens=0)
    assert causal.decoder_causal_mask is True


def test_learnable_pos_emb_is_rejected():
    from egom2p_b200 import adapters as A
    with pytest.raises(NotImplementedError, match="pos_emb"):
        e.create_model("egom2p_tiny_6e_6d_swiglu_nobias",
                       encoder_embeddings={"tok_cam": A.GazeCamTokenEncoderEmbedding(sincos_pos_emb=False)},
                       decoder_embeddings={"tok_cam": A.GazeCamTokenDecoderEmbedding()},
                       modality_info={"tok_cam": MI["tok_cam"]}, num_register_tokens=0)
