"""Host-side generation schedules (egom2p_b200/generate.py) against the unmodified reference's
build_chained_generation_schedules / linear_schedule / cosine_schedule (egom2p/models/generate.py:130-321), pinned by
tests/golden/schedules_ref.json (oracle/gen_golden_schedules.py ran the live reference): the four eval workloads plus chained,
MaskGIT-cosine, linear / onex temperature and no-grow variants."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gen_golden_schedules as ggs  # noqa: E402  (the CASES table only; the reference is not imported here)


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "schedules_ref.json")))


@pytest.mark.parametrize("case", list(ggs.CASES))
def test_chained_schedule_equals_reference(golden, case):
    from egom2p_b200.generate import build_chained_generation_schedules
    ours = build_chained_generation_schedules(**ggs.CASES[case])
    ref = golden["schedules"][case]
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert a["target_domain"] == b["target_domain"] and a["scheme"] == b["scheme"]
        assert int(a["num_tokens"]) == b["num_tokens"]                       # token counts: exact
        assert list(a["cfg_cond_domains"]) == b["cfg_cond_domains"]
        assert float(a["cfg_scale"]) == b["cfg_scale"]
        assert float(a["temperature"]) == pytest.approx(b["temperature"], rel=1e-12, abs=0.0)
    assert sum(int(a["num_tokens"]) for a in ours) == sum(ggs.CASES[case]["tokens_per_target"])


def test_token_schedules_equal_reference(golden):
    from egom2p_b200.generate import cosine_schedule, linear_schedule
    for key, ref in golden["helpers"]["linear_schedule"].items():
        steps, total = map(int, key.split(","))
        assert [int(v) for v in linear_schedule(steps, total)] == ref
    for key, ref in golden["helpers"]["cosine_schedule"].items():
        steps, total = map(int, key.split(","))
        assert [int(v) for v in cosine_schedule(steps, total)] == ref


def test_illegal_schemes_raise():
    from egom2p_b200.generate import build_chained_generation_schedules
    kw = dict(ggs.CASES["rgb2cam"])
    with pytest.raises(ValueError):
        build_chained_generation_schedules(**{**kw, "autoregression_schemes": ["bogus"]})
    with pytest.raises(ValueError):
        build_chained_generation_schedules(**{**kw, "temp_schedules": ["bogus"]})
    with pytest.raises(NotImplementedError):
        build_chained_generation_schedules(**{**kw, "autoregression_schemes": ["autoregressive"]})
