"""TEST INFRASTRUCTURE ONLY. Full-size parity fixtures for the configuration bench.py times (BASELINE.json configs[1]):
the UNMODIFIED reference `egom2p_base_12e_12d_swiglu_nobias` (created through the reference registry,
egom2p_model.py:1053-1074), fp32 on CPU, b = 1, N = M = 2048, on two batches:

  * "dense":  the SURVEY 8(d) headline split (1009 rgb + 1009 depth + 15 cam + 15 gaze inputs and as many targets);
  * "ragged": sample 0 of tests/golden/ref_masks_egob.npz (masks drawn by the reference UnifiedMasking).

Weights = oracle/synth.make_state_dict(ego-b cfg, seed 0) (regenerated identically on the GPU box: numpy PCG64).
Stores loss, per-modality loss, every parameter's gradient norm and sum, a few whole gradients, and per head the
logits of that modality's valid target rows: logsumexp over the whole vocabulary plus a strided 16-column slice for
the 64k heads, the full 256 columns for the cam / gaze heads (where the 2e-2 bound is tight).
Run (authoring container only, ~3 min of CPU): `python oracle/gen_golden_egob.py`."""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402
import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MODS = ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"]
N_ENC = N_DEC = 2048
SD_SEED = 0
SHUFFLE_SEED = 7
COLS = np.arange(16) * 3989 + 11   # strided vocabulary columns of the 64k heads (max 59846)


def egob_cfg():
    return synth.make_cfg(768, 12, 12, 12, MODS)


def dense_batch(cfg):
    n = {"tok_cam": [15], "tok_depth": [1009], "tok_gaze": [15], "tok_rgb": [1009]}
    return synth.make_batch(cfg, B=1, seed=41, n_in=n, n_tgt=n)


def ragged_batch(cfg, index=0):
    g = np.load(os.path.join(OUT, "ref_masks_egob.npz"))
    rng = np.random.default_rng(43)
    md = {}
    for m, inf in cfg["mods"].items():
        L = inf["len"]
        t = torch.from_numpy(rng.integers(0, inf["vocab"], size=(1, L), dtype=np.int64))
        if "thw" in inf:
            t = t.reshape(1, *inf["thw"])
        md[m] = {"tensor": t,
                 "input_mask": torch.from_numpy(np.unpackbits(g[m + "_input_mask"][index])[:L].astype(bool)[None]),
                 "target_mask": torch.from_numpy(np.unpackbits(g[m + "_target_mask"][index])[:L].astype(bool)[None]),
                 "decoder_attention_mask": torch.from_numpy(g[m + "_attn"][index].astype(np.int32)[None])}
    return md


def clone_md(md):
    return {m: {k: v.clone() for k, v in d.items()} for m, d in md.items()}


SAVE_GRADS = ("mask_token", "decoder_proj_context.bias", "encoder_norm.weight", "decoder_norm.weight",
              "encoder.0.norm1.weight", "encoder.11.norm2.weight", "decoder.0.query_norm.weight", "decoder.11.context_norm.weight",
              "encoder_embeddings.tok_cam.token_emb.weight", "encoder_embeddings.tok_rgb.mod_emb",
              "decoder_embeddings.tok_gaze.token_emb.weight")


def main():
    import_reference()
    from egom2p.data.modality_info import MODALITY_INFO
    from egom2p.utils.timm.model_builder import create_model
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = egob_cfg()
    enc = {m: MODALITY_INFO[m]["encoder_embedding"]() for m in MODS}
    dec = {m: MODALITY_INFO[m]["decoder_embedding"]() for m in MODS}
    model = create_model("egom2p_base_12e_12d_swiglu_nobias", encoder_embeddings=enc, decoder_embeddings=dec,
                         modality_info={m: MODALITY_INFO[m] for m in MODS}, num_register_tokens=0)
    model.load_state_dict(synth.make_state_dict(cfg, SD_SEED), strict=True)
    for name, md in (("dense", dense_batch(cfg)), ("ragged", ragged_batch(cfg))):
        res = {"cols": COLS}
        random.seed(SHUFFLE_SEED)
        res["dec_order"] = np.array(random.sample(MODS, len(MODS)))
        model.zero_grad()
        random.seed(SHUFFLE_SEED)
        loss, mod_loss = model(clone_md(md), N_ENC, N_DEC, loss_type="mod")
        loss.backward()
        res["loss"] = loss.detach().double().numpy()
        res["mod_loss"] = np.array([mod_loss[m].item() for m in MODS], dtype=np.float64)
        names, norms, sums = [], [], []
        for n, p in model.named_parameters():
            names.append(n)
            norms.append(p.grad.double().norm().item())
            sums.append(p.grad.double().sum().item())
            if n in SAVE_GRADS:
                res["grad::" + n] = p.grad.numpy().copy()
        res["grad_names"], res["grad_norms"], res["grad_sums"] = np.array(names), np.array(norms), np.array(sums)
        model.zero_grad()
        with torch.no_grad():
            random.seed(SHUFFLE_SEED)
            logits = model(clone_md(md), N_ENC, N_DEC, return_logits=True)
            # rows of each modality in the compacted decoder sequence: recover them the way the reference lays them out
            dec_d = {m: model.decoder_embeddings[m].forward_embed(d) for m, d in clone_md(md).items()}
            random.seed(SHUFFLE_SEED)
            _, _, dmask, tgt, _, dmod = model.forward_mask_decoder(dec_d, N_DEC)
        res["dec_mod"] = dmod.numpy()
        res["target_ids"] = tgt.numpy()
        for m in MODS:
            rows = (dmod[0] == MODALITY_INFO[m]["id"]).nonzero().reshape(-1)
            lg = logits[m][0, rows]
            res[f"rows::{m}"] = rows.numpy().astype(np.int32)
            res[f"lse::{m}"] = torch.logsumexp(lg.double(), -1).numpy()
            res[f"logits::{m}"] = lg.numpy().copy() if lg.shape[-1] <= 256 else lg[:, torch.from_numpy(COLS)].numpy().copy()
        del logits
        np.savez_compressed(os.path.join(OUT, f"egob_{name}.npz"), **res)
        print(name, "loss", float(res["loss"]), res["mod_loss"], "order", list(res["dec_order"]),
              "rows", {m: len(res[f"rows::{m}"]) for m in MODS}, flush=True)


if __name__ == "__main__":
    main()
