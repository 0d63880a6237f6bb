"""tcgen05 GEMM vs torch fp32 matmul on bf16-rounded inputs (all operand major-ness combos)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(torch.bfloat16).cuda()


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 512, 512), (6400, 1536, 512), (200, 80, 512),
                                   (1024, 512, 1536), (333, 200, 80)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    from kokoro_ruslan_b200 import ops
    if a_mn and M % 8:
        pytest.skip("MN-major A needs M % 8 == 0 (TMA stride)")
    if b_mn and N % 8:
        pytest.skip("MN-major B needs N % 8 == 0")
    A = _mk((M, K), 1)
    B = _mk((N, K), 2)
    ref = A.float() @ B.float().t()
    a_arg = A.t().contiguous() if a_mn else A
    b_arg = B.t().contiguous() if b_mn else B
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a_arg, b_arg, out, a_mn_major=a_mn, b_mn_major=b_mn)
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err


def test_gemm_epilogues():
    from kokoro_ruslan_b200 import ops
    M, N, K = 640, 384, 512
    A, B = _mk((M, K), 3), _mk((N, K), 4)
    bias = torch.randn(N, device="cuda")
    resid = torch.randn(M, N, device="cuda")
    ref = A.float() @ B.float().t()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, B, out, bias=bias)
    assert torch.allclose(out.float(), ref + bias, atol=0.1, rtol=2e-2)
    out32 = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, out32, bias=bias, resid=resid, alpha=0.5)
    assert torch.allclose(out32, 0.5 * ref + bias + resid, atol=2e-2, rtol=2e-3)
    # row-modulo residual (positional table broadcast over the batch)
    pe = torch.randn(64, N, device="cuda")
    ops.gemm(A, B, out32, resid=pe, resid_mod=64)
    assert torch.allclose(out32, ref + pe.repeat(M // 64, 1), atol=2e-2, rtol=2e-3)
    # split-K atomic accumulate on top of existing content
    acc = torch.ones(M, N, device="cuda")
    ops.gemm(A, B, acc, accumulate=True, splits=4)
    assert torch.allclose(acc, ref + 1.0, atol=2e-2, rtol=2e-3)


def test_gemm_batched():
    from kokoro_ruslan_b200 import ops
    A, B = _mk((5, 200, 128), 5), _mk((5, 96, 128), 6)
    out = torch.empty(5, 200, 96, device="cuda")
    ops.gemm(A, B, out)
    ref = torch.einsum("bmk,bnk->bmn", A.float(), B.float())
    assert torch.allclose(out, ref, atol=2e-2, rtol=2e-3)
