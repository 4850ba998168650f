"""GPU: fc1|fc3 GEMM with the SwiGLU gate fused into its epilogue, and the fc2 dgrad with the SwiGLU derivative fused,
against a torch fp32 evaluation on the same bf16 operands (tolerance: bf16 output rounding, 2e-2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def interleave(w1, w3):
    F, D = w1.shape
    return torch.stack([w1.view(F // 32, 32, D), w3.view(F // 32, 32, D)], 1).reshape(2 * F, D)


@pytest.mark.parametrize("R,D,F", [(300, 256, 704), (4096, 768, 2048), (77, 384, 1024)])
def test_swiglu_gemm_fwd_bwd(R, D, F):
    from egom2p_b200 import ops
    gen = torch.Generator().manual_seed(R)
    x = torch.randn(R, D, generator=gen).bfloat16()
    w1 = (torch.randn(F, D, generator=gen) / D ** 0.5).bfloat16()
    w3 = (torch.randn(F, D, generator=gen) / D ** 0.5).bfloat16()
    w2 = (torch.randn(D, F, generator=gen) / F ** 0.5).bfloat16()
    dy = torch.randn(R, D, generator=gen).bfloat16()
    a = (x.float() @ w1.float().t())
    b = (x.float() @ w3.float().t())
    g_ref = torch.nn.functional.silu(a) * b
    ab, g = ops.gemm_swiglu_fwd(x.cuda(), interleave(w1, w3).cuda())
    abv = ab.float().cpu().view(R, F // 32, 2, 32)
    torch.testing.assert_close(abv[:, :, 0].reshape(R, F), a, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(abv[:, :, 1].reshape(R, F), b, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(g.float().cpu(), g_ref, rtol=2e-2, atol=2e-2)
    # backward: dg = dy @ w2 ; da = dg * b * silu'(a) ; db = dg * silu(a), with a, b as stored (bf16)
    a_s, b_s = abv[:, :, 0].reshape(R, F), abv[:, :, 1].reshape(R, F)
    dg = dy.float() @ w2.float()
    sg = torch.sigmoid(a_s)
    da_ref = dg * b_s * (sg * (1 + a_s * (1 - sg)))
    db_ref = dg * a_s * sg
    dab = ops.gemm_swiglu_bwd(dy.cuda(), w2.cuda(), ab).float().cpu().view(R, F // 32, 2, 32)
    torch.testing.assert_close(dab[:, :, 0].reshape(R, F), da_ref, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(dab[:, :, 1].reshape(R, F), db_ref, rtol=3e-2, atol=3e-2)
