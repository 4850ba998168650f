"""TEST INFRASTRUCTURE ONLY. Pins the host-side generation schedules (egom2p_b200/generate.py) against the UNMODIFIED
reference (egom2p/models/generate.py:197-321 build_chained_generation_schedules and the token / temperature schedule helpers,
imported from /root/reference): the four eval workloads (eval_model_rgb2depth.py:45-59, eval_model_rgb2cam.py:40-54,
eval_model_rgb2gaze.py:41-55, eval_model_depth2rgb.py:34-48) plus chained / MaskGIT / temperature-schedule variants.
-> tests/golden/schedules_ref.json. Run: `python oracle/gen_golden_schedules.py`."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "schedules_ref.json")

CASES = {
    "rgb2depth": dict(cond_domains=["tok_rgb"], target_domains=["tok_depth"], tokens_per_target=[5120], autoregression_schemes=["roar"],
                      decoding_steps=[3], token_decoding_schedules=["linear"], temps=[0.01], temp_schedules=["constant"],
                      cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True),
    "rgb2cam": dict(cond_domains=["tok_rgb"], target_domains=["tok_cam"], tokens_per_target=[30], autoregression_schemes=["roar"],
                    decoding_steps=[3], token_decoding_schedules=["linear"], temps=[0.01], temp_schedules=["constant"],
                    cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True),
    "rgb2gaze": dict(cond_domains=["tok_rgb"], target_domains=["tok_gaze"], tokens_per_target=[30], autoregression_schemes=["roar"],
                     decoding_steps=[5], token_decoding_schedules=["linear"], temps=[0.01], temp_schedules=["constant"],
                     cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True),
    "depth2rgb": dict(cond_domains=["tok_depth"], target_domains=["tok_rgb"], tokens_per_target=[5120], autoregression_schemes=["roar"],
                      decoding_steps=[6], token_decoding_schedules=["linear"], temps=[0.01], temp_schedules=["constant"],
                      cfg_scales=[2.0], cfg_schedules=["constant"], cfg_grow_conditioning=True),
    "chained_maskgit_cosine_linear_temp": dict(
        cond_domains=["tok_rgb"], target_domains=["tok_depth", "tok_cam", "tok_gaze"], tokens_per_target=[5120, 30, 30],
        autoregression_schemes=["maskgit", "roar", "maskgit"], decoding_steps=[7, 4, 30], token_decoding_schedules=["cosine", "linear", "linear"],
        temps=[1.5, 0.7, 3.0], temp_schedules=["linear", "constant", "onex:0.5:0.5"], cfg_scales=[1.0, 2.5, 0.0],
        cfg_schedules=["constant"] * 3, cfg_grow_conditioning=True),
    "no_grow_uneven": dict(
        cond_domains=["tok_depth", "tok_gaze"], target_domains=["tok_rgb", "tok_cam"], tokens_per_target=[5119, 29],
        autoregression_schemes=["roar", "roar"], decoding_steps=[11, 29], token_decoding_schedules=["linear", "linear"],
        temps=[0.2, 1.0], temp_schedules=["onex:0.1:2.0", "linear"], cfg_scales=[3.0, 1.0], cfg_schedules=["constant"] * 2,
        cfg_grow_conditioning=False),
}


def main():
    import_reference()
    from egom2p.models import generate as ref
    out = {}
    for name, kw in CASES.items():
        sched = ref.build_chained_generation_schedules(**kw)
        out[name] = [{"target_domain": s["target_domain"], "scheme": s["scheme"], "num_tokens": int(s["num_tokens"]),
                      "temperature": float(s["temperature"]), "cfg_scale": float(s["cfg_scale"]),
                      "cfg_cond_domains": list(s["cfg_cond_domains"])} for s in sched]
    helpers = {"linear_schedule": {}, "cosine_schedule": {}}
    for steps, total in [(3, 5120), (6, 5120), (5, 30), (7, 30), (30, 30), (50, 30), (1, 17), (13, 1000)]:
        helpers["linear_schedule"][f"{steps},{total}"] = [int(v) for v in ref.linear_schedule(steps, total)]
        helpers["cosine_schedule"][f"{steps},{total}"] = [int(v) for v in ref.cosine_schedule(steps, total)]
    json.dump({"schedules": out, "helpers": helpers}, open(OUT, "w"), indent=1)
    print("wrote", OUT, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
