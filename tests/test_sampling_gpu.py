"""GPU: fused token sampling (egom2p_sample_rows) and classifier-free-guidance combine against the CPU restatement of
the reference's sample_tokens / top_k_top_p_filtering (oracle/sampling_oracle.py, pinned by the live reference's filter):
the set of tokens that survive top-k / top-p, the drawn token for the same uniforms (inverse CDF in token order), its
probability, argmax at temperature 0 -- on the 256-entry cam / gaze vocabulary and the 64000-entry Cosmos codebook."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gen_golden_sampling as ggs  # noqa: E402
import sampling_oracle as so  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from egom2p_b200 import ops as o
    return o


@pytest.mark.parametrize("case", ggs.CASES, ids=[c[0] for c in ggs.CASES])
def test_sample_rows_matches_oracle(ops, case):
    name, rows, V, scale, top_k, top_p, temp = case
    lg = ggs.make_logits(name, rows, V, scale)
    rng = np.random.default_rng(7)
    reps = 8
    lg_rep = np.repeat(lg, reps, axis=0)
    u = rng.random(rows * reps).astype(np.float32)
    k_int = (min(top_k, V) if isinstance(top_k, int) else min(int(top_k * V), V)) if top_k > 0 else 0
    tok, prob, kept = ops.sample_rows(torch.from_numpy(lg_rep).cuda(), temp, top_p, k_int, torch.from_numpy(u).cuda(), want_kept=True)
    tok, prob, kept = tok.cpu().numpy(), prob.cpu().numpy(), kept.cpu().numpy()
    keep = so.kept_mask(lg, top_k, top_p)
    assert np.abs(kept.reshape(rows, reps)[:, 0] - keep.sum(1)).max() <= 1          # filtered set (boundary token +-1)
    want_tok, want_prob, margin = so.draw(lg_rep, temp, u.astype(np.float64), top_k, top_p)
    solid = margin > 1e-4                                                            # draws not sitting on a CDF step
    assert solid.mean() > 0.9
    assert np.array_equal(tok[solid], want_tok[solid])
    np.testing.assert_allclose(prob[solid], want_prob[solid], rtol=2e-3, atol=1e-6)
    assert keep[np.repeat(np.arange(rows), reps), tok].all()                         # never a filtered-out token


def test_sample_rows_greedy_and_distribution(ops):
    g = torch.Generator().manual_seed(0)
    lg = torch.randn(5, 64000, generator=g) * 3
    lg[2, 777] = lg[2].max() + 1
    lg[2, 12345] = lg[2, 777]                                                        # tie: torch.argmax returns the first
    tok, prob, _ = ops.sample_rows(lg.cuda(), 0.0, 0.8, 0)
    assert torch.equal(tok.cpu(), lg.argmax(-1)) and bool((prob == 1).all())
    # distribution at temperature 1, no filter: empirical frequencies of a 256-way categorical follow softmax(logits)
    lg = torch.randn(1, 256, generator=g) * 1.5
    n = 200000
    tok, _, _ = ops.sample_rows(lg.cuda().expand(n, 256).contiguous(), 1.0, 0.0, 0)
    freq = torch.bincount(tok.cpu(), minlength=256).double() / n
    p = torch.softmax(lg[0].double(), -1)
    assert float((freq - p).abs().max()) < 4 * float((p.max() * (1 - p.max()) / n) ** 0.5) + 1e-3


def test_cfg_combine_equals_logit_space_guidance(ops):
    """W (y_u + s (y_c - y_u)) == l_u + s (l_c - l_u) (generate.py:804) up to the bf16 rounding of the GEMM operand."""
    g = torch.Generator().manual_seed(1)
    yu, yc = torch.randn(300, 768, generator=g), torch.randn(300, 768, generator=g)
    w = (torch.randn(512, 768, generator=g) * 0.02)
    comb = ops.cfg_combine_bf16(yu.cuda(), yc.cuda(), 2.0)
    ref = yu + (yc - yu) * 2.0
    assert torch.equal(comb.cpu(), ref.bfloat16())
    lu, lc = yu.double() @ w.double().t(), yc.double() @ w.double().t()
    want = lu + (lc - lu) * 2.0
    got = ops.linear_fwd(comb, ops.cast_bf16(w.cuda()), out_dtype=torch.float32).cpu().double()
    assert float((got - want).abs().max()) < 2e-2
