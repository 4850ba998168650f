"""GPU: the one-launch operand refresh (egom2p_cast_f32_to_bf16_multi) is bit-exact with a plain round-to-nearest-even cast,
for plain, padded-pitch, interleaved (fc1 | fc3) and odd-width (scalar path) destinations."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_cast_plan_bit_exact():
    from egom2p_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(3)
    bf16 = torch.bfloat16

    def rnd(r, c):
        return (torch.randn(r, c, generator=g) * 3).to(dev)

    a = rnd(300, 768)            # plain
    b = rnd(64, 682)             # odd width -> scalar path, destination pitch padded to 704
    c1, c3 = rnd(100, 256), rnd(100, 256)   # interleaved pair, 100 rows padded to 128
    d = rnd(1, 8)                # single tiny row
    e = rnd(9000, 64)            # many chunks
    da = torch.zeros(300, 768, dtype=bf16, device=dev)
    db = torch.zeros(64, 704, dtype=bf16, device=dev)
    dc = torch.full((256, 256), 7.0, dtype=bf16, device=dev)
    dd = torch.zeros(1, 8, dtype=bf16, device=dev)
    de = torch.zeros(9000, 64, dtype=bf16, device=dev)
    plan = ops.CastPlan([(a, da, 0, 0), (b, db, 0, 0), (c1, dc, 32, 0), (c3, dc, 32, 1), (d, dd, 0, 0), (e, de, 0, 0)])
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(da, a.to(bf16))
    assert torch.equal(db[:, :682], b.to(bf16)) and float(db[:, 682:].abs().max()) == 0.0
    v = dc.view(4, 2, 32, 256)
    assert torch.equal(v[:, 0].reshape(128, 256)[:100], c1.to(bf16))
    assert torch.equal(v[:, 1].reshape(128, 256)[:100], c3.to(bf16))
    assert torch.equal(v[3, :, 4:], torch.full((2, 28, 256), 7.0, dtype=bf16, device=dev))   # rows past 100 untouched
    assert torch.equal(dd, d.to(bf16))
    assert torch.equal(de, e.to(bf16))
