"""The two fused sub-layer tails of the forward chain — dropout + residual (+ RMSNorm) AND the LayerNorm of the sub-layer
that follows in one row-wise kernel (kr_resid_drop_ln_fwd, kr_rmsnorm_resid_ln_fwd) — against (a) the pair of kernels they
replace, under the same dropout masks, and (b) a plain torch fp32 statement of the same rows.  Reference semantics:
kokoro/model/transformers.py (pre-norm residual blocks; x = x + dropout(sublayer(norm(x)))) — the fusion only changes which
kernel writes what.  Tolerances: the sum x + y is exact either way; the statistics may differ by FMA contraction (<= 2e-6 rel
on fp32), bf16 outputs by at most one bf16 ulp on a few elements."""
import pytest
import torch

pytestmark = pytest.mark.gpu

M, D, S = 6400, 512, 800


def _inputs():
    g = torch.Generator().manual_seed(2)
    a = (torch.randn(M, D, generator=g) * 0.3).to(torch.bfloat16).cuda()
    w = (torch.randn(D, D, generator=g) / D ** 0.5).to(torch.bfloat16).cuda()
    bias, x = torch.randn(D, generator=g).cuda(), torch.randn(M, D, generator=g).cuda()
    gam, bet = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    rs = (torch.rand(M // S, generator=g) + 0.5).cuda()
    return a, w, bias, x, gam, bet, rs


def _specs(rs):
    from kokoro_ruslan_b200 import ops
    state = torch.tensor([7, 9], dtype=torch.int64, device="cuda")
    return (("none", None), ("one mask", ops.make_drop_spec(state, 3, 0.2)),
            ("two masks + row scale", ops.make_drop_spec(state, 3, 0.2, 4, 0.1, row_scale=rs, rows_per_sample=S)))


def _ln_torch(out):
    mean = out.mean(-1)
    var = out.var(-1, unbiased=False)
    return mean, torch.rsqrt(var + 1e-5)


@pytest.mark.parametrize("f32_out", [False, True])
def test_out_projection_tail_with_following_layernorm(f32_out):
    from kokoro_ruslan_b200 import ops
    a, w, bias, x, gam, bet, rs = _inputs()
    for name, spec in _specs(rs):
        out1 = torch.empty(M, D, device="cuda")
        m1, r1 = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
        h1 = torch.empty(M, D, dtype=torch.float32 if f32_out else torch.bfloat16, device="cuda")
        ops.gemm(a, w, out1, bias=bias, resid=x, drop=spec)                    # round-1 path: epilogue does dropout + residual
        ops.layernorm_fwd(out1, gam, bet, None if f32_out else h1, h1 if f32_out else None, m1, r1)
        yo, out2, h2 = torch.empty(M, D, device="cuda"), torch.empty(M, D, device="cuda"), torch.empty_like(h1)
        m2, r2 = torch.empty(M, device="cuda"), torch.empty(M, device="cuda")
        ops.gemm(a, w, yo, bias=bias)
        ops.resid_drop_ln_fwd(yo, x, out2, spec, gam, bet, None if f32_out else h2, h2 if f32_out else None, m2, r2)
        torch.cuda.synchronize()
        assert float((out1 - out2).abs().max()) <= 2e-6 * float(out1.abs().max()), name      # same masks, same sum
        if spec is None:
            assert torch.equal(out2, yo + x), name
        mt, rt = _ln_torch(out2)
        assert torch.allclose(m2, mt, atol=2e-6, rtol=1e-5) and torch.allclose(r2, rt, rtol=2e-5), name
        assert torch.allclose(m2, m1, atol=2e-6, rtol=1e-5) and torch.allclose(r2, r1, rtol=2e-5), name
        ht = (out2 - mt[:, None]) * rt[:, None] * gam + bet
        tol = 1e-5 if f32_out else 2 ** -7                                      # fp32 output | one bf16 ulp
        assert float((h2.float() - ht).abs().max()) <= tol * float(ht.abs().max()), name
        assert float((h2.float() - h1.float()).abs().max()) <= tol * float(ht.abs().max()), name


def test_ffn_tail_with_following_layernorm():
    from kokoro_ruslan_b200 import ops
    _, _, _, x, gam, bet, rs = _inputs()
    g = torch.Generator().manual_seed(5)
    y, gain = torch.randn(M, D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    for name, spec in _specs(rs):
        o1, o2 = torch.empty(M, D, device="cuda"), torch.empty(M, D, device="cuda")
        h1 = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
        h2 = torch.empty_like(h1)
        m1, r1, m2, r2 = (torch.empty(M, device="cuda") for _ in range(4))
        ops.rmsnorm_resid_fwd(y, gain, x, o1, drop=spec)
        ops.layernorm_fwd(o1, gam, bet, h1, None, m1, r1)
        ops.rmsnorm_resid_ln_fwd(y, gain, x, o2, spec, gam, bet, h2, None, m2, r2)
        torch.cuda.synchronize()
        assert float((o1 - o2).abs().max()) <= 2e-6 * float(o1.abs().max()), name
        if spec is None:
            rms = y * torch.rsqrt((y * y).mean(-1, keepdim=True) + torch.finfo(torch.float32).eps) * gain
            assert float((o2 - (x + rms)).abs().max()) <= 1e-5 * float(o2.abs().max()), name
        mt, rt = _ln_torch(o2)
        assert torch.allclose(m2, mt, atol=2e-6, rtol=1e-5) and torch.allclose(r2, rt, rtol=2e-5), name
        ht = (o2 - mt[:, None]) * rt[:, None] * gam + bet
        assert float((h2.float() - ht).abs().max()) <= 2 ** -7 * float(ht.abs().max()), name
        assert float((h2.float() - h1.float()).abs().max()) <= 2 ** -7 * float(ht.abs().max()), name
