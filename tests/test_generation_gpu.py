"""GPU, T3 of SURVEY.md section 7: generation passes of the B200 sampler (egom2p_b200/generate.py) at the real sizes of
BASELINE.json configs[2..3] against the UNMODIFIED reference GenerationSampler.forward_enc_dec_roar_batched in fp32
(tests/golden/generation_egob.npz, oracle/gen_golden_generation.py): ego-b weights, encoder N in {0, 10, 3414, 4267, 5120, 5130,
8534, 9387}, decoder k in {10, 853, 1706}, conditional and unconditional branch batched in one decoder pass. Plus: batching ragged
samples equals running them one by one, and end-to-end `generate()` invariants on a small model."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import synth  # noqa: E402
import gen_golden_egob as gg  # noqa: E402
import gen_golden_generation as ggen  # noqa: E402


@pytest.fixture(scope="module")
def egob_sampler():
    import egom2p_b200 as e
    from egom2p_b200.generate import GenerationSampler
    from egom2p_b200.modality_info import MODALITY_INFO as MI
    model = e.create_model("egom2p_base_12e_12d_swiglu_nobias",
                           encoder_embeddings={k: MI[k]["encoder_embedding"]() for k in gg.MODS},
                           decoder_embeddings={k: MI[k]["decoder_embedding"]() for k in gg.MODS},
                           modality_info={k: MI[k] for k in gg.MODS}, num_register_tokens=0)
    model.load_state_dict(synth.make_state_dict(gg.egob_cfg(), gg.SD_SEED), strict=True)
    return GenerationSampler(model.cuda().eval())


@pytest.mark.parametrize("case", list(ggen.CASES))
def test_guided_roar_pass_matches_reference(egob_sampler, golden_dir, case):
    from egom2p_b200 import ops
    g = np.load(os.path.join(golden_dir, "generation_egob.npz"))
    s = egob_sampler
    target, n_done, k = ggen.CASES[case]
    md = ggen.make_state(case, "cuda")
    pos = torch.from_numpy(g[f"{case}::pos"]).cuda()           # the positions the reference drew (CPU generator)
    n_in = {m: [int((~d["input_mask"]).sum())] for m, d in md.items()}
    yn = s.forward_hidden(md, target, [ggen.COND[case]], n_in, pos)   # (2, k, D): conditional, unconditional
    wb = s.model.head_operand(target)
    cols = torch.from_numpy(g["cols"]).cuda()
    worst = {}
    for i, name in enumerate(("cond", "uncond")):
        lg = ops.linear_fwd(ops.cast_bf16(yn[i].contiguous()), wb, out_dtype=torch.float32)
        got = lg if lg.shape[-1] <= 256 else lg.index_select(1, cols)
        worst[name] = float((got - torch.from_numpy(g[f"{case}::{name}::logits"]).cuda()).abs().max())
        worst[name + "/lse"] = float(np.abs(torch.logsumexp(lg.double(), -1).cpu().numpy() - g[f"{case}::{name}::lse"]).max())
    assert all(v < 2e-2 for v in worst.values()), worst
    # fused guidance + head + greedy sampling == argmax of the reference's guided logits wherever that argmax is not a near-tie
    yb = ops.cfg_combine_bf16(yn[1].contiguous(), yn[0].contiguous(), ggen.SCALE)
    tok, _ = s.sample_from_hidden(yb, target, 0.0, 0.0, 0.8)
    clear = g[f"{case}::guided_gap"] > 0.1
    assert np.array_equal(tok.cpu().numpy()[clear], g[f"{case}::guided_argmax"][clear]), (case, int(clear.sum()))


def _small_sampler():
    from egom2p_b200.generate import GenerationSampler
    from test_model_gpu import build_model
    cfg = synth.make_cfg(192, 3, 2, 2, ["tok_cam", "tok_depth", "tok_gaze", "tok_rgb"], video_vocab=512, video_thw=(5, 4, 4))
    model = build_model(cfg).cuda().eval()
    model.load_state_dict(synth.make_state_dict(cfg, 3), strict=True)
    return GenerationSampler(model), cfg


def test_ragged_batch_equals_single_samples():
    """Two samples with different numbers of conditioning tokens in one batch == each sample alone (ranges, not padding, decide
    what is attended), for the conditional and the unconditional branch."""
    s, cfg = _small_sampler()
    rng = np.random.default_rng(0)
    B, L = 2, 80
    rgb_in = np.ones((B, L), dtype=bool)
    rgb_in[0, :] = False
    rgb_in[1, rng.permutation(L)[:37]] = False
    md = {"tok_rgb": {"tensor": torch.from_numpy(rng.integers(0, 512, (B, 5, 4, 4))).cuda(), "input_mask": torch.from_numpy(rgb_in).cuda(),
                      "target_mask": torch.ones(B, L, dtype=torch.bool).cuda()},
          "tok_cam": {"tensor": torch.from_numpy(rng.integers(0, 256, (B, 30))).cuda(), "input_mask": torch.ones(B, 30, dtype=torch.bool).cuda(),
                      "target_mask": torch.zeros(B, 30, dtype=torch.bool).cuda()}}
    md["tok_cam"]["input_mask"][:, :4] = False
    md["tok_cam"]["target_mask"][:, :4] = True
    pos = torch.tensor([[7, 11, 29, 5, 20]] * B).cuda()
    n_in = {"tok_rgb": [80, 37], "tok_cam": [4, 4]}
    both = s.forward_hidden(md, "tok_cam", ["tok_rgb"], n_in, pos).reshape(2, B, 5, -1)
    for b in range(B):
        one = {m: {k: v[b:b + 1] for k, v in d.items()} for m, d in md.items()}
        alone = s.forward_hidden(one, "tok_cam", ["tok_rgb"], {m: [v[b]] for m, v in n_in.items()}, pos[b:b + 1]).reshape(2, 5, -1)
        assert float((alone - both[:, b]).abs().max()) < 2e-2 * float(alone.abs().max())


@pytest.mark.parametrize("scheme", ["roar", "maskgit"])
def test_generate_end_to_end_small(scheme):
    from egom2p_b200.generate import build_chained_generation_schedules, init_empty_target_modality, init_full_input_modality
    s, cfg = _small_sampler()
    info = s.model.modality_info
    B = 3
    rng = np.random.default_rng(1)
    md = {"tok_rgb": {"tensor": torch.from_numpy(rng.integers(0, 512, (B, 5, 4, 4))).cuda()}}
    md = init_empty_target_modality(md, info, "tok_cam", B, 30, "cuda")
    md = init_empty_target_modality(md, info, "tok_depth", B, 80, "cuda")
    md = init_full_input_modality(md, info, "tok_rgb", "cuda")
    schedule = build_chained_generation_schedules(
        cond_domains=["tok_rgb"], target_domains=["tok_cam", "tok_depth"], tokens_per_target=[30, 80], autoregression_schemes=[scheme] * 2,
        decoding_steps=[3, 4], token_decoding_schedules=["linear"] * 2, temps=[0.01, 1.0], temp_schedules=["constant", "linear"],
        cfg_scales=[2.0, 1.5], cfg_schedules=["constant"] * 2, cfg_grow_conditioning=True)
    assert [st["num_tokens"] for st in schedule] == [10, 10, 10, 20, 20, 20, 20]
    rgb_before = md["tok_rgb"]["tensor"].clone()
    out = s.generate(md, schedule, top_p=0.8, top_k=0.0, seed=0)
    again = s.generate(md, schedule, top_p=0.8, top_k=0.0, seed=0)
    for mod, V in (("tok_cam", 256), ("tok_depth", 512)):
        assert bool(out[mod]["target_mask"].all()) and not bool(out[mod]["input_mask"].any())   # every position decoded
        assert int(out[mod]["tensor"].min()) >= 0 and int(out[mod]["tensor"].max()) < V
        assert int(md[mod]["tensor"].abs().sum()) == 0                                            # the caller's dict is untouched
    assert torch.equal(out["tok_rgb"]["tensor"], rgb_before)
    assert torch.equal(out["tok_cam"]["tensor"], again["tok_cam"]["tensor"])                       # same seed, greedy-like temperature
    # 7 guided steps per run: the conditional encoder pass every step; the unconditional one is skipped while its context is
    # empty (first cam step; first depth step, where rgb AND the finished cam are conditioning): 7 + 5 passes per run
    assert s.stats["steps"] == 14 and s.stats["encoder_passes"] == 2 * (7 + 5)


def test_graphed_generation_matches_eager_greedy():
    """GraphedGeneration (whole generate() call as one CUDA graph) decodes the same tokens as the eager call at temperature 0
    when both use the same decoding order: MaskGIT positions are deterministic, so the two runs must agree exactly."""
    from egom2p_b200.generate import (GraphedGeneration, build_chained_generation_schedules, init_empty_target_modality,
                                      init_full_input_modality)
    s, cfg = _small_sampler()
    info = s.model.modality_info
    B = 2
    rng = np.random.default_rng(5)

    def make(seed):
        md = {"tok_rgb": {"tensor": torch.from_numpy(np.random.default_rng(seed).integers(0, 512, (B, 5, 4, 4))).cuda()}}
        md = init_empty_target_modality(md, info, "tok_depth", B, 80, "cuda")
        return init_full_input_modality(md, info, "tok_rgb", "cuda")
    schedule = build_chained_generation_schedules(
        cond_domains=["tok_rgb"], target_domains=["tok_depth"], tokens_per_target=[80], autoregression_schemes=["maskgit"],
        decoding_steps=[4], token_decoding_schedules=["linear"], temps=[0.0], temp_schedules=["constant"],
        cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True)
    graphed = GraphedGeneration(s, make(0), schedule, top_p=0.8)
    for seed in (1, 2):
        md = make(seed)
        want = s.generate(md, schedule, top_p=0.8, seed=None)["tok_depth"]["tensor"]
        got = graphed(md)["tok_depth"]["tensor"]
        assert bool(graphed.static_out["tok_depth"]["target_mask"].all())
        # greedy decoding: identical up to argmax near-ties between two separately scheduled runs (none expected: same kernels)
        assert float((got == want).float().mean()) > 0.98
