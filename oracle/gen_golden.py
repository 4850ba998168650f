"""TEST INFRASTRUCTURE ONLY. Runs the UNMODIFIED reference (imported from /root/reference, authoring
container only) on deterministic inputs from oracle/synth.py and writes its outputs to tests/golden/.
Run: `python oracle/gen_golden.py`. The fixtures pin oracle/egom2p_oracle.py (tests/test_oracle.py).
"""
import json
import os
import random
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)

ref_model = import_reference()
from egom2p.data.modality_info import MODALITY_INFO  # noqa: E402
from egom2p.models.egom2p_utils import LayerNorm  # noqa: E402
from egom2p.models import encoder_embeddings as ee, decoder_embeddings as de  # noqa: E402


def build_ref(cfg, share_embedding=True):
    enc, dec = {}, {}
    for m, inf in cfg["mods"].items():
        if "thw" in inf:
            isz = inf["thw"][1] * 8
            enc[m] = ee.VideoTokenEncoderEmbedding(vocab_size=inf["vocab"], image_size=isz)
            dec[m] = de.VideoTokenDecoderEmbedding(vocab_size=inf["vocab"], image_size=isz, share_embedding=share_embedding)
        else:
            enc[m] = ee.GazeCamTokenEncoderEmbedding(vocab_size=inf["vocab"])
            dec[m] = de.GazeCamTokenDecoderEmbedding(vocab_size=inf["vocab"], share_embedding=share_embedding)
    return ref_model.EgoM2P(
        encoder_embeddings=enc, decoder_embeddings=dec, modality_info={m: MODALITY_INFO[m] for m in cfg["mods"]},
        dim=cfg["dim"], encoder_depth=cfg["enc_depth"], decoder_depth=cfg["dec_depth"], num_heads=cfg["heads"],
        mlp_ratio=4, qkv_bias=False, proj_bias=False, mlp_bias=False,
        norm_layer=partial(LayerNorm, eps=1e-6, bias=False), act_layer=torch.nn.SiLU, gated_mlp=True)


def clone_md(md):
    return {m: {k: v.clone() for k, v in d.items()} for m, d in md.items()}


def dec_order_for(seed, mods):
    random.seed(seed)
    return [m for m in random.sample(list(mods), len(mods))]


def run_case(name, cfg, sd_seed, batch_kw, n_enc, n_dec, shuffle_seed, tie=True):
    model = build_ref(cfg, share_embedding=tie)
    sd = synth.make_state_dict(cfg, sd_seed, tie=tie)
    missing = model.load_state_dict(sd, strict=True)
    md = synth.make_batch(cfg, **batch_kw)
    order = dec_order_for(shuffle_seed, list(cfg["mods"]))
    res = {"dec_order": np.array(order)}

    # index plan + gathered embeddings straight from the reference's own methods
    with torch.no_grad():
        m1 = clone_md(md)
        enc_d = {m: model.encoder_embeddings[m](d) for m, d in m1.items()}
        et, ee_, em, emod = model.forward_mask_encoder(enc_d, n_enc)
        m2 = clone_md(md)
        dec_d = {m: model.decoder_embeddings[m].forward_embed(d) for m, d in m2.items()}
        random.seed(shuffle_seed)
        dt, demb, dm, tgt, dattn, dmod = model.forward_mask_decoder(dec_d, n_dec)
    res.update(enc_x0=(et + ee_).numpy(), enc_emb=ee_.numpy(), enc_mask=em[:, 0].numpy(), enc_mod=emod.numpy(),
               dec_y0=(dt + demb).numpy(), dec_mask=dm[:, 0].numpy(), dec_mod=dmod.numpy(), target_ids=tgt.numpy(),
               dec_attn=np.packbits(dattn.numpy(), axis=-1))

    for lt in ("mod", "token"):
        model.zero_grad()
        random.seed(shuffle_seed)
        loss, mod_loss = model(clone_md(md), n_enc, n_dec, loss_type=lt)
        res[f"loss_{lt}"] = loss.detach().numpy()
        res[f"mod_loss_{lt}"] = np.array([mod_loss[m].item() for m in cfg["mods"]], dtype=np.float64)
        if lt == "mod":
            loss.backward()
            names, norms, sums = [], [], []
            for n, p in model.named_parameters():
                names.append(n)
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                norms.append(g.double().norm().item())
                sums.append(g.double().sum().item())
            res["grad_names"] = np.array(names)
            res["grad_norms"] = np.array(norms)
            res["grad_sums"] = np.array(sums)
            for n, p in model.named_parameters():
                if n in ("mask_token", "decoder_proj_context.bias", "encoder_norm.weight", "decoder_norm.weight",
                         "encoder.0.attn.qkv.weight", "decoder.1.cross_attn.kv.weight",
                         "encoder_embeddings.tok_cam.token_emb.weight", "encoder_embeddings.tok_cam.mod_emb",
                         "decoder_embeddings.tok_gaze.token_emb.weight", "decoder.0.mlp.fc2.weight"):
                    res["grad::" + n] = p.grad.numpy()
    with torch.no_grad():
        random.seed(shuffle_seed)
        logits = model(clone_md(md), n_enc, n_dec, return_logits=True)
    for m in cfg["mods"]:
        lg = logits[m].numpy()
        res[f"logits_head::{m}"] = lg[:, :, :16].copy()      # first 16 vocab columns, all rows
        res[f"logits_lse::{m}"] = torch.logsumexp(logits[m].double(), -1).numpy()
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **res)
    print(name, "loss", float(res["loss_mod"]), res["mod_loss_mod"], "order", order)


def main():
    torch.manual_seed(0)
    # ---- small 4-modality model (T1-style semantics at toy width), ragged budgets incl. empty modalities
    cfg = synth.make_cfg(48, 2, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=128,
                         video_thw=(5, 4, 4))
    bk = dict(B=4, seed=11,
              n_in={"tok_cam": [5, 0, 30, 2], "tok_depth": [30, 10, 0, 1], "tok_gaze": [4, 0, 30, 0], "tok_rgb": [25, 40, 4, 0]},
              n_tgt={"tok_cam": [10, 30, 0, 1], "tok_depth": [20, 0, 40, 0], "tok_gaze": [3, 0, 0, 0], "tok_rgb": [15, 18, 70, 0]})
    run_case("small4", cfg, 5, bk, 64, 48, shuffle_seed=3)
    # ---- BASELINE config c1: tiny FM (dim 256, 2+2 layers, 4 heads) on example_data/token cam+gaze tokens
    cfg1 = synth.make_cfg(256, 4, 2, 2, ["tok_cam", "tok_gaze"])
    model = ref_model.FM({"domains_in": ["tok_cam", "tok_gaze"], "domains_out": ["tok_cam", "tok_gaze"], "dim": 256,
                          "encoder_depth": 2, "decoder_depth": 2, "num_heads": 4, "mlp_ratio": 4.0, "qkv_bias": False,
                          "proj_bias": False, "mlp_bias": False, "norm_bias": False, "act_layer": "SiLU",
                          "gated_mlp": True, "image_size": 256, "patch_size": 8})
    sd = synth.make_state_dict(cfg1, 7, tie=False)
    model.load_state_dict(sd, strict=True)
    cam = np.load("/root/reference/example_data/token/cam-tok.npz")
    gaze = np.load("/root/reference/example_data/token/gaze-tok.npz")
    cam_ids = cam[cam.files[0]].reshape(1, 30).astype(np.int64)
    gaze_ids = gaze[gaze.files[0]].reshape(1, 30).astype(np.int64)
    md = synth.make_batch(cfg1, B=1, seed=21, n_in={"tok_cam": [14], "tok_gaze": [10]}, n_tgt={"tok_cam": [8], "tok_gaze": [12]})
    md["tok_cam"]["tensor"] = torch.from_numpy(cam_ids)
    md["tok_gaze"]["tensor"] = torch.from_numpy(gaze_ids)
    order = dec_order_for(1, ["tok_cam", "tok_gaze"])
    random.seed(1)
    loss, mod_loss = model(clone_md(md), 24, 20)
    loss.backward()
    res = {"cam_ids": cam_ids, "gaze_ids": gaze_ids, "dec_order": np.array(order), "loss": loss.detach().numpy(),
           "mod_loss": np.array([mod_loss[m].item() for m in ["tok_cam", "tok_gaze"]]),
           "grad_names": np.array([n for n, _ in model.named_parameters()]),
           "grad_norms": np.array([0.0 if p.grad is None else p.grad.double().norm().item() for _, p in model.named_parameters()])}
    np.savez_compressed(os.path.join(OUT, "c1_tiny_fm.npz"), **res)
    print("c1", float(loss), res["mod_loss"])

    # ---- ego-b shaped index plan (B=4, L=10300, budgets 2048/2048), ragged + dense rows
    cfgp = synth.make_cfg(12, 1, 0, 0, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=64000)
    model = build_ref(cfgp)
    n_in = {"tok_cam": [15, 0, 30, 0], "tok_depth": [1009, 694, 0, 30], "tok_gaze": [15, 0, 30, 30], "tok_rgb": [1009, 0, 5120, 0]}
    n_tg = {"tok_cam": [15, 30, 0, 1], "tok_depth": [1009, 0, 2048, 0], "tok_gaze": [15, 0, 0, 0], "tok_rgb": [1009, 2018, 0, 0]}
    md = synth.make_batch(cfgp, B=4, seed=31, n_in=n_in, n_tgt=n_tg)
    with torch.no_grad():
        enc_d = {m: model.encoder_embeddings[m](d) for m, d in clone_md(md).items()}
        _, _, em, emod = model.forward_mask_encoder(enc_d, 2048)
        dec_d = {m: model.decoder_embeddings[m].forward_embed(d) for m, d in clone_md(md).items()}
        random.seed(5)
        _, _, dm, tgt, dattn, dmod = model.forward_mask_decoder(dec_d, 2048)
        # recover ids_keep the way the reference computes it (egom2p_model.py:370-373)
        _, _, mask_all, _ = model.cat_encoder_tensors(enc_d)
        ar = torch.arange(mask_all.shape[1]).unsqueeze(0) * 1e-6
        enc_keep = torch.argsort(mask_all + ar, dim=1)[:, :2048]
    order = dec_order_for(5, list(cfgp["mods"]))
    np.savez_compressed(os.path.join(OUT, "plan_egob.npz"), dec_order=np.array(order), enc_keep=enc_keep.numpy().astype(np.int32),
                        enc_mask=em[:, 0].numpy(), enc_mod=emod.numpy(), dec_mask=dm[:, 0].numpy(), dec_mod=dmod.numpy(),
                        target_ids=tgt.numpy(), dec_attn=np.packbits(dattn.numpy(), axis=-1))
    print("plan_egob ok", order)

    # ---- ego-b state_dict manifest straight from the reference registry
    from egom2p.utils.timm.model_builder import create_model  # noqa
    mods = ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"]
    with torch.device("meta"):
        pass
    enc = {m: MODALITY_INFO[m]["encoder_embedding"]() for m in mods}
    dec = {m: MODALITY_INFO[m]["decoder_embedding"]() for m in mods}
    model = create_model("egom2p_base_12e_12d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec,
                         modality_info={m: MODALITY_INFO[m] for m in mods}, num_register_tokens=0)
    man = {k: list(v.shape) for k, v in model.state_dict().items()}
    params = [n for n, _ in model.named_parameters()]
    with open(os.path.join(OUT, "egob_state_dict_manifest.json"), "w") as f:
        json.dump({"state_dict": man, "named_parameters": params,
                   "n_params": sum(p.numel() for p in model.parameters())}, f, indent=0)
    print("manifest", len(man), len(params), sum(p.numel() for p in model.parameters()))


if __name__ == "__main__":
    main()
