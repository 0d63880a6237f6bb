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


@pytest.mark.parametrize("block_n", [0, 64, 128, 192, 256])
def test_gemm_block_n_variants_persistent(block_n):
    """Every tile width, many tiles per CTA (persistent loop + double-buffered TMEM accumulator)."""
    import ctypes
    from kokoro_ruslan_b200 import ops
    from kokoro_ruslan_b200._lib import check, lib
    M, N, K = 128 * 41 + 77, 1000, 704
    A, B = _mk((M, K), 11), _mk((N, K), 12)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    a = ops.GemmArgs()
    a.A, a.B, a.M, a.N, a.K, a.batch = A.data_ptr(), B.data_ptr(), M, N, K, 1
    a.lda, a.ldb, a.alpha, a.beta, a.bias = K, K, 1.0, 1.0, bias.data_ptr()
    a.C, a.c_mode, a.ldc, a.splits, a.force_block_n = out.data_ptr(), 1, N, 1, block_n
    check(lib().kr_gemm_ex(ctypes.byref(a), ops._stream()), "kr_gemm_ex")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t() + bias
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err


@pytest.mark.parametrize("cin,cout,k,dil", [(128, 128, 3, 1), (128, 128, 7, 3), (64, 64, 11, 5), (512, 256, 7, 1),
                                            (64, 200, 3, 1), (32, 32, 3, 1), (32, 32, 7, 3), (32, 32, 11, 5), (32, 64, 7, 1)])
def test_conv1d_channels_last_implicit_gemm(cin, cout, k, dil):
    """conv mode of the GEMM kernel vs torch conv1d (same bf16-rounded operands), with the fused
    bias + bf16 residual + 1/3 scaling + second leaky-relu output used by the HiFi-GAN MRF."""
    from kokoro_ruslan_b200 import ops
    Bn, L, halo = 3, 1000, 32
    pad = dil * (k - 1) // 2
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(Bn, cin, L, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g)
    res = (torch.randn(Bn, cout, L, generator=g)).to(torch.bfloat16)
    acc = (torch.randn(Bn, cout, L, generator=g)).to(torch.bfloat16)
    ref = torch.nn.functional.conv1d(x.float(), w.float(), bias, padding=pad, dilation=dil) + res.float()
    ref2 = ref / 3.0 + acc.float()
    xcl = torch.zeros(Bn, L + 2 * halo, cin, dtype=torch.bfloat16, device="cuda")
    xcl[:, halo:halo + L] = x.permute(0, 2, 1).cuda()
    wt = w.permute(0, 2, 1).reshape(cout, k * cin).contiguous().cuda()          # tap-major K
    out = torch.zeros(Bn, L + 2 * halo, cout, dtype=torch.bfloat16, device="cuda")
    act = torch.zeros_like(out)
    rcl = res.permute(0, 2, 1).contiguous().cuda()
    acl = acc.permute(0, 2, 1).contiguous().cuda()
    ops.conv1d_cl(xcl, wt, rows=L, row0=halo - pad, taps=k, dil=dil, bias=bias.cuda(), out=out[:, halo:halo + L],
                  out_act=act[:, halo:halo + L], act_slope=0.1, resid=rcl, resid2=acl, beta=1.0 / 3.0)
    torch.cuda.synchronize()
    got = out[:, halo:halo + L].float().cpu().permute(0, 2, 1)
    scale = ref2.abs().max().item()
    assert (got - ref2).abs().max().item() / scale < 1e-2
    got_act = act[:, halo:halo + L].float().cpu().permute(0, 2, 1)
    assert (got_act - torch.nn.functional.leaky_relu(ref2, 0.1)).abs().max().item() / scale < 1e-2
    assert float(out[:, :halo].abs().max()) == 0.0 and float(out[:, halo + L:].abs().max()) == 0.0   # halos untouched


@pytest.mark.parametrize("cin,cout,k,dil", [(64, 64, 3, 1), (64, 64, 7, 3), (64, 64, 11, 5), (128, 128, 3, 1), (64, 128, 3, 1),
                                            (32, 32, 3, 1), (32, 32, 7, 3), (32, 32, 11, 5)])
def test_conv1d_slab_mode_matches_tap_refetch(cin, cout, k, dil):
    """Large-M conv (enough tiles for the resident-weight + slab path): one activation slab per tile with the
    taps as row-shifted UMMA descriptors must give the same result as re-fetching every tap, and match torch."""
    from kokoro_ruslan_b200 import ops
    Bn, L, halo = 4, 128 * 90 + 37, 32
    pad = dil * (k - 1) // 2
    g = torch.Generator().manual_seed(9)
    x = (torch.randn(Bn, cin, L, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g)
    ref = torch.nn.functional.conv1d(x.float(), w.float(), bias, padding=pad, dilation=dil)
    xcl = torch.zeros(Bn, L + 2 * halo, cin, dtype=torch.bfloat16, device="cuda")
    xcl[:, halo:halo + L] = x.permute(0, 2, 1).cuda()
    wt = w.permute(0, 2, 1).reshape(cout, k * cin).contiguous().cuda()
    outs = []
    for no_slab in (True, False):
        out = torch.zeros(Bn, L, cout, dtype=torch.float32, device="cuda")
        ops.conv1d_cl(xcl, wt, rows=L, row0=halo - pad, taps=k, dil=dil, bias=bias.cuda(), out=out, no_slab=no_slab)
        torch.cuda.synchronize()
        outs.append(out.cpu().permute(0, 2, 1))
    scale = ref.abs().max().item()
    assert (outs[0] - ref).abs().max().item() / scale < 5e-3
    assert (outs[1] - ref).abs().max().item() / scale < 5e-3, "slab mode differs from torch"
    assert torch.equal(outs[0], outs[1]), "slab mode is not bit-identical to the tap re-fetch path"


@pytest.mark.parametrize("cin,cout,k,dil,Bn,L", [(128, 128, 7, 3, 3, 256 * 60 + 37), (128, 128, 11, 5, 3, 256 * 60 + 200),
                                                  (256, 256, 3, 1, 2, 128 * 90 + 5), (256, 256, 11, 5, 2, 128 * 80),
                                                  (64, 64, 11, 1, 4, 256 * 50 + 129), (64, 64, 11, 5, 1, 700),
                                                  (512, 256, 7, 1, 1, 1000)])
def test_conv1d_slab_stream_mode(cin, cout, k, dil, Bn, L):
    """Convs whose weights do NOT fit in shared memory (the 128 / 256-channel HiFi-GAN stages, k = 11 of the 64-channel
    one): activation slab per (tile, 64-channel block), streamed weight blocks, two 128-row sub-tiles per block where TMEM
    allows.  Against torch and against the tap re-fetch path (same products, different fp32 summation order), including
    the fused residual / MRF / activation epilogue on ragged row counts."""
    from kokoro_ruslan_b200 import ops
    halo = 32
    pad = dil * (k - 1) // 2
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(Bn, cin, L, generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, generator=g) / (cin * k) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g)
    resid = torch.randn(Bn, L, cout, generator=g)
    resid2 = torch.randn(Bn, L, cout, generator=g)
    conv = torch.nn.functional.conv1d(x.float(), w.float(), bias, padding=pad, dilation=dil).permute(0, 2, 1)
    want = (conv + resid) * (1.0 / 3.0) + resid2
    xcl = torch.zeros(Bn, L + 2 * halo, cin, dtype=torch.bfloat16, device="cuda")
    xcl[:, halo:halo + L] = x.permute(0, 2, 1).cuda()
    wt = w.permute(0, 2, 1).reshape(cout, k * cin).contiguous().cuda()
    outs, acts = [], []
    for no_slab in (True, False):
        out = torch.zeros(Bn, L, cout, dtype=torch.float32, device="cuda")
        act = torch.zeros(Bn, L + 2 * halo, cout, dtype=torch.bfloat16, device="cuda")
        ops.conv1d_cl(xcl, wt, rows=L, row0=halo - pad, taps=k, dil=dil, bias=bias.cuda(), out=out,
                      out_act=act[:, halo:halo + L], act_slope=0.1, resid=resid.cuda(), resid2=resid2.cuda(), beta=1.0 / 3.0,
                      no_slab=no_slab)
        torch.cuda.synchronize()
        outs.append(out.cpu())
        acts.append(act.float().cpu())
    scale = want.abs().max().item()
    assert (outs[1] - want).abs().max().item() / scale < 5e-3
    assert (outs[1] - outs[0]).abs().max().item() / scale < 1e-5
    assert float(acts[1][:, :halo].abs().max()) == 0.0 and float(acts[1][:, halo + L:].abs().max()) == 0.0
    lre = torch.nn.functional.leaky_relu(outs[1], 0.1)
    assert (acts[1][:, halo:halo + L] - lre).abs().max().item() / scale < 1e-2
