"""Thin Python marshalling over the C ABI in include/kokoro_b200.h.

Every function takes torch CUDA tensors, passes raw device pointers + the current CUDA stream to
libkokoro_b200.so and raises RuntimeError on a non-zero return code.  No torch math happens here.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from ._lib import check, lib

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float

EPI_BF16, EPI_F32, EPI_ATOMIC_F32 = 0, 1, 2


class DropSpec(ctypes.Structure):
    """ctypes mirror of kr_drop_spec (include/kokoro_b200.h): one dropout / stochastic-depth site."""
    _fields_ = [("state", c_void_p), ("site_a", ctypes.c_uint), ("thr_a", ctypes.c_uint),
                ("site_b", ctypes.c_uint), ("thr_b", ctypes.c_uint), ("scale", c_float),
                ("row_scale", c_void_p), ("rows_per_sample", c_int)]


def drop_thr(p: float) -> int:
    """16-bit keep threshold of drop probability p: keep iff lane16 >= thr."""
    return max(0, min(65535, int(round(float(p) * 65536.0))))


def drop_keep(p: float) -> float:
    """Exact keep probability realised by drop_thr(p)."""
    return 1.0 - drop_thr(p) / 65536.0


def drop_thr8(p: float) -> int:
    """Threshold of the byte-lane mode (attention-probability sites): p quantised to 1/256, in 16-bit units."""
    return max(0, min(255, int(round(float(p) * 256.0)))) * 256


def make_drop_spec(state: torch.Tensor, site_a: int = 0, p_a: float = 0.0, site_b: int = 0, p_b: float = 0.0,
                   row_scale: Optional[torch.Tensor] = None, rows_per_sample: int = 1,
                   byte_lanes: bool = False) -> Optional["DropSpec"]:
    """None when the site is a no-op (both probabilities 0 and no per-sample factors).  byte_lanes: the
    attention kernels' generator mode (threshold quantised to 1/256; the scale stays the exact 1 / keep)."""
    ta, tb = (drop_thr8(p_a) if byte_lanes else drop_thr(p_a)), drop_thr(p_b)
    if ta == 0 and tb == 0 and row_scale is None:
        return None
    if ta == 0 and tb != 0:
        site_a, ta, site_b, tb = site_b, tb, 0, 0
    d = DropSpec()
    d.state = state.data_ptr()
    d.site_a, d.thr_a, d.site_b, d.thr_b = site_a, ta, site_b, tb
    d.scale = 1.0 / ((1.0 - ta / 65536.0) * (1.0 - tb / 65536.0))
    d.row_scale = row_scale.data_ptr() if row_scale is not None else None
    d.rows_per_sample = rows_per_sample
    d._keep = (state, row_scale)       # the spec borrows device memory
    return d


def _dref(d: Optional["DropSpec"]):
    return ctypes.byref(d) if d is not None else None


def _ptr(t: Optional[torch.Tensor]) -> c_void_p:
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RuntimeError("kokoro_ruslan_b200 ops need CUDA tensors (no CPU fallback)")
    return c_void_p(t.data_ptr())


def _stream() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, a_mn_major: bool = False,
         b_mn_major: bool = False, bias: Optional[torch.Tensor] = None,
         resid: Optional[torch.Tensor] = None, resid_mod: int = 0, alpha: float = 1.0,
         accumulate: bool = False, splits: int = 1, drop: Optional[DropSpec] = None, block_n: int = 0) -> torch.Tensor:
    """out[M,N] (+)= alpha * A @ B^T (+bias) (+resid) on tcgen05 (bf16 in, fp32 accumulate).

    A logical [M,K]: stored [M,K] (K-major) or, if a_mn_major, stored [K,M].
    B logical [N,K]: stored [N,K] (K-major) or, if b_mn_major, stored [K,N].
    2-D or batched 3-D (leading batch dim) operands; `out` dtype selects the epilogue
    (bf16 / fp32); accumulate=True uses fp32 atomics (split-K allowed).
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    batched = a.dim() == 3
    if batched:
        batch = a.shape[0]
        a2, b2, o2 = a[0], b[0], out[0]
        sa, sb, sc = a.stride(0), b.stride(0), out.stride(0)
    else:
        batch, a2, b2, o2, sa, sb, sc = 1, a, b, out, 0, 0, 0
    assert a2.stride(1) == 1 and b2.stride(1) == 1 and o2.stride(1) == 1
    if a_mn_major:
        K, M = a2.shape
    else:
        M, K = a2.shape
    if b_mn_major:
        Kb, N = b2.shape
    else:
        N, Kb = b2.shape
    assert K == Kb, (a.shape, b.shape)
    assert tuple(o2.shape) == (M, N), (o2.shape, M, N)
    if accumulate:
        assert out.dtype == torch.float32
        epi = EPI_ATOMIC_F32
    else:
        epi = EPI_BF16 if out.dtype == torch.bfloat16 else EPI_F32
        assert out.dtype in (torch.bfloat16, torch.float32)
    ldr, sr = 0, 0
    if resid is not None:
        assert resid.dtype == torch.float32
        r2 = resid[0] if (batched and resid.dim() == 3) else resid
        ldr = r2.stride(0)
        sr = resid.stride(0) if (batched and resid.dim() == 3) else 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if drop is not None:             # dropout epilogue: v = drop(alpha*acc + bias) + resid  (kr_gemm_ex only)
        assert not batched and not accumulate and splits == 1
        g = GemmArgs()
        g.A, g.B, g.M, g.N, g.K, g.batch = _p(a), _p(b), M, N, K, 1
        g.lda, g.ldb = a2.stride(0), b2.stride(0)
        g.a_mn_major, g.b_mn_major = int(a_mn_major), int(b_mn_major)
        g.alpha, g.beta = alpha, 1.0
        g.bias = _p(bias)
        if resid is not None:
            g.resid, g.resid_dtype, g.ldr, g.resid_mod = _p(resid), 0, ldr, resid_mod
        g.C, g.c_mode, g.ldc = _p(out), epi, o2.stride(0)
        g.splits = 1
        g.force_block_n = block_n
        g.drop = ctypes.addressof(drop)
        check(lib().kr_gemm_ex(ctypes.byref(g), _stream()), "kr_gemm_ex")
        return out
    rc = lib().kr_gemm_bf16(_ptr(a), _ptr(b), _ptr(out), c_int(M), c_int(N), c_int(K), c_int(batch),
                            c_ll(a2.stride(0)), c_ll(b2.stride(0)), c_ll(o2.stride(0)),
                            c_ll(sa), c_ll(sb), c_ll(sc), c_int(int(a_mn_major)),
                            c_int(int(b_mn_major)), c_int(epi), _ptr(bias), _ptr(resid),
                            c_ll(ldr), c_ll(sr), c_int(resid_mod), c_float(alpha), c_int(splits),
                            _stream())
    check(rc, "kr_gemm_bf16")
    return out


class GemmArgs(ctypes.Structure):
    """ctypes mirror of kr_gemm_args (include/kokoro_b200.h)."""
    _fields_ = [("A", c_void_p), ("B", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int), ("batch", c_int),
                ("lda", c_ll), ("ldb", c_ll), ("stride_a", c_ll), ("stride_b", c_ll),
                ("a_mn_major", c_int), ("b_mn_major", c_int),
                ("conv_taps", c_int), ("conv_dil", c_int), ("conv_row0", c_int), ("conv_cin", c_int),
                ("a_rows", c_int),
                ("alpha", c_float), ("beta", c_float),
                ("bias", c_void_p),
                ("resid", c_void_p), ("resid_dtype", c_int), ("ldr", c_ll), ("stride_r", c_ll), ("resid_mod", c_int),
                ("resid2", c_void_p), ("resid2_dtype", c_int), ("ldr2", c_ll), ("stride_r2", c_ll),
                ("C", c_void_p), ("c_mode", c_int), ("ldc", c_ll), ("stride_c", c_ll),
                ("C2", c_void_p), ("ldc2", c_ll), ("stride_c2", c_ll), ("act_slope", c_float),
                ("splits", c_int), ("force_block_n", c_int), ("no_slab", c_int), ("drop", c_void_p)]


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("kokoro_ruslan_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def conv1d_cl(x: torch.Tensor, w: torch.Tensor, *, rows: int, row0: int, taps: int, dil: int,
              bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
              out_act: Optional[torch.Tensor] = None, act_slope: float = 0.1,
              resid: Optional[torch.Tensor] = None, resid2: Optional[torch.Tensor] = None,
              beta: float = 1.0, block_n: int = 0, no_slab: bool = False) -> None:
    """Implicit-GEMM conv1d on channels-last bf16 activations (tcgen05).

    x: [B, rows_phys, C_in] bf16 with zero halos; output row m (0 <= m < rows) of item b reads
    x[b, row0 + m + tap*dil, :] for tap in range(taps).  w: [N, taps*C_in] bf16 (tap-major K).
    out / out_act / resid / resid2: [B, rows, N]-shaped *views* (any row / batch strides, unit
    inner stride): v = acc + bias + resid; v = v*beta + resid2; out = v; out_act = lrelu(v).
    """
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and x.dim() == 3 and w.dim() == 2
    Bn, rows_phys, cin = x.shape
    N, K = w.shape
    assert K == taps * cin and x.stride(2) == 1 and w.stride(1) == 1
    a = GemmArgs()
    a.A, a.B = _p(x), _p(w)
    a.M, a.N, a.K, a.batch = rows, N, K, Bn
    a.lda, a.ldb, a.stride_a, a.stride_b = x.stride(1), w.stride(0), x.stride(0), 0
    a.conv_taps, a.conv_dil, a.conv_row0, a.conv_cin, a.a_rows = taps, dil, row0, cin, rows_phys
    a.alpha, a.beta = 1.0, beta
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        a.bias = _p(bias)

    def _view(t):
        assert t.dim() == 3 and t.shape[0] == Bn and t.shape[1] >= rows and t.shape[2] == N and t.stride(2) == 1, \
            (t.shape, t.stride(), rows, N)
        return _p(t), t.stride(1), t.stride(0)
    if resid is not None:
        a.resid, a.ldr, a.stride_r = _view(resid)
        a.resid_dtype = int(resid.dtype == torch.bfloat16)
    if resid2 is not None:
        a.resid2, a.ldr2, a.stride_r2 = _view(resid2)
        a.resid2_dtype = int(resid2.dtype == torch.bfloat16)
    if out is not None:
        a.C, a.ldc, a.stride_c = _view(out)
        a.c_mode = EPI_BF16 if out.dtype == torch.bfloat16 else EPI_F32
    else:
        a.c_mode = 3
    if out_act is not None:
        assert out_act.dtype == torch.bfloat16
        a.C2, a.ldc2, a.stride_c2 = _view(out_act)
        a.act_slope = act_slope
    a.splits, a.force_block_n, a.no_slab = 1, block_n, int(no_slab)
    check(lib().kr_gemm_ex(ctypes.byref(a), _stream()), "kr_gemm_ex")


def hifi_resblock(x_act: torch.Tensor, L: int, halo: int, w1: torch.Tensor, off1, kh1, b1: torch.Tensor,
                  w2: torch.Tensor, off2, kh2, b2: torch.Tensor, resid: torch.Tensor, *, resid2: Optional[torch.Tensor] = None,
                  beta: float = 1.0, out: Optional[torch.Tensor] = None, out_act: Optional[torch.Tensor] = None,
                  act_slope: float = 0.1) -> None:
    """One fused HiFi-GAN ResBlock step (csrc/kr_hifi_resblock.cu): v = conv2(lrelu(conv1(x_act) + b1)) + b2 + resid;
    v = v * beta + resid2; out = v; out_act = lrelu(v, act_slope).  x_act: the padded channels-last activation
    [B, L + 2*halo, 64] bf16 (zero halos); both convs as K-half-block lists (hifigan.HiFiGANGenerator._half_blocks):
    w [64, n*32] bf16, row offsets off[n] and input-channel halves kh[n] (host ints); resid / resid2 / out / out_act:
    [B, L, 64]-shaped views (any row / batch strides, unit inner stride) that start at time 0."""
    assert x_act.dtype == torch.bfloat16 and x_act.dim() == 3 and x_act.is_contiguous()
    B, rows_phys, C = x_act.shape
    n1, n2 = len(off1), len(off2)
    assert C == 64 and rows_phys == L + 2 * halo and w1.shape == (C, n1 * 32) and w2.shape == (C, n2 * 32)
    assert len(kh1) == n1 and len(kh2) == n2
    assert w1.is_contiguous() and w2.is_contiguous() and w1.dtype == torch.bfloat16 and w2.dtype == torch.bfloat16
    arr = lambda v: (c_int * len(v))(*[int(t) for t in v])     # noqa: E731

    def view(t, dtype):
        if t is None:
            return None, 0, 0
        assert t.dtype == dtype and tuple(t.shape) == (B, L, C) and t.stride(2) == 1, (t.shape, t.stride(), (B, L, C))
        return t, t.stride(1), t.stride(0)
    r, r_ld, r_bs = view(resid, torch.float32)
    r2, r2_ld, r2_bs = view(resid2, torch.float32)
    o, o_ld, o_bs = view(out, torch.float32)
    a, a_ld, a_bs = view(out_act, torch.bfloat16)
    check(lib().kr_hifi_resblock(_ptr(x_act), c_int(B), c_ll(L), c_int(halo), _ptr(w1), c_int(n1), arr(off1), arr(kh1), _ptr(b1),
                                 _ptr(w2), c_int(n2), arr(off2), arr(kh2), _ptr(b2), _ptr(r), c_ll(r_ld), c_ll(r_bs), _ptr(r2),
                                 c_ll(r2_ld), c_ll(r2_bs), c_float(beta), _ptr(o), c_ll(o_ld), c_ll(o_bs), _ptr(a), c_ll(a_ld),
                                 c_ll(a_bs), c_float(act_slope), _stream()), "kr_hifi_resblock")


def _heads_strides(t: torch.Tensor):
    """t is a [B, S, H, 64] bf16 view with unit inner stride and head stride 64."""
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.shape[3] == 64
    assert t.stride(3) == 1 and t.stride(2) == 64, t.stride()
    return c_ll(t.stride(1)), c_ll(t.stride(0))


def attn_fwd(q, k, v, o, lse, key_mask: Optional[torch.Tensor], causal: bool, scale: float,
             drop: Optional[DropSpec] = None):
    """Flash attention forward (tcgen05).  q,o: [B,Sq,H,64]; k,v: [B,Sk,H,64]; lse: [B,H,Sq] f32
    (log2 domain); key_mask: [B,Sk] uint8 (1 = masked) or None."""
    B, Sq, H, _ = q.shape
    Sk = k.shape[1]
    if key_mask is not None:
        assert key_mask.dtype == torch.uint8 and key_mask.is_contiguous() and key_mask.shape == (B, Sk)
    assert lse.dtype == torch.float32 and lse.is_contiguous()
    rc = lib().kr_attn_fwd(_ptr(q), *_heads_strides(q), _ptr(k), *_heads_strides(k), _ptr(v),
                           *_heads_strides(v), _ptr(o), *_heads_strides(o), _ptr(lse),
                           _ptr(key_mask), c_int(B), c_int(H), c_int(Sq), c_int(Sk),
                           c_int(int(causal)), c_float(scale), _dref(drop), _stream())
    check(rc, "kr_attn_fwd")


def attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, key_mask, causal: bool, scale: float,
             drop: Optional[DropSpec] = None):
    """Flash attention backward.  dq: fp32 [B,Sq,H,64] (zeroed by the call's own prep kernel, then atomically
    accumulated); dk, dv: bf16 [B,Sk,H,64]; delta: fp32 scratch [B,H,Sq]."""
    B, Sq, H, _ = q.shape
    Sk = k.shape[1]
    assert dq.dtype == torch.float32 and dq.stride(3) == 1 and dq.stride(2) == 64
    rc = lib().kr_attn_bwd(_ptr(q), *_heads_strides(q), _ptr(k), *_heads_strides(k), _ptr(v),
                           *_heads_strides(v), _ptr(o), *_heads_strides(o), _ptr(d_o),
                           *_heads_strides(d_o), _ptr(lse), _ptr(delta), _ptr(dq),
                           c_ll(dq.stride(1)), c_ll(dq.stride(0)), _ptr(dk), *_heads_strides(dk),
                           _ptr(dv), *_heads_strides(dv), _ptr(key_mask), c_int(B), c_int(H),
                           c_int(Sq), c_int(Sk), c_int(int(causal)), c_float(scale), _dref(drop), _stream())
    check(rc, "kr_attn_bwd")


# ---------------------------------------------------------------------------------------------
# HBM kernels
# ---------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, y_bf16, y_f32, mean, rstd, eps: float = 1e-5):
    N, D = x.shape
    check(lib().kr_layernorm_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(y_bf16), _ptr(y_f32), _ptr(mean),
                                 _ptr(rstd), c_int(N), c_int(D), c_float(eps), _stream()), "kr_layernorm_fwd")


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, dx_bf16, dgamma, dbeta, drop_bf16: Optional[DropSpec] = None,
                  dcol_bf16: Optional[torch.Tensor] = None):
    """drop_bf16: dropout / stochastic-depth factors applied to the bf16 copy only (it feeds the backward of
    the dropped residual branch that precedes this norm; dx itself is the residual-stream gradient).
    dcol_bf16 [D] += column sums of that bf16 copy: the bias gradient of the Linear whose backward it enters."""
    N, D = x.shape
    check(lib().kr_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres), _ptr(dx),
                                 _ptr(dx_bf16), _ptr(dgamma), _ptr(dbeta), c_int(N), c_int(D), _dref(drop_bf16),
                                 _ptr(dcol_bf16), _stream()),
          "kr_layernorm_bwd")


def rmsnorm_resid_fwd(y, gain, resid, out, drop: Optional[DropSpec] = None):
    N, D = y.shape
    check(lib().kr_rmsnorm_resid_fwd(_ptr(y), _ptr(gain), _ptr(resid), _ptr(out), c_int(N), c_int(D), _dref(drop),
                                     _stream()),
          "kr_rmsnorm_resid_fwd")


def resid_drop_ln_fwd(y, resid, out, drop: Optional[DropSpec], gamma, beta, h_bf16, h_f32, mean, rstd, eps: float = 1e-5):
    """out = resid + dropout(y); h = LayerNorm(out) (csrc/kr_norm.cu resid_drop_ln_fwd_kernel)."""
    N, D = y.shape
    check(lib().kr_resid_drop_ln_fwd(_ptr(y), _ptr(resid), _ptr(out), c_int(N), c_int(D), _dref(drop), _ptr(gamma), _ptr(beta),
                                     _ptr(h_bf16), _ptr(h_f32), _ptr(mean), _ptr(rstd), c_float(eps), _stream()),
          "kr_resid_drop_ln_fwd")


def rmsnorm_resid_ln_fwd(y, gain, resid, out, drop: Optional[DropSpec], gamma, beta, h_bf16, h_f32, mean, rstd,
                         eps: float = 1e-5):
    """rmsnorm_resid_fwd + the LayerNorm that follows it, one kernel."""
    N, D = y.shape
    check(lib().kr_rmsnorm_resid_ln_fwd(_ptr(y), _ptr(gain), _ptr(resid), _ptr(out), c_int(N), c_int(D), _dref(drop), _ptr(gamma),
                                        _ptr(beta), _ptr(h_bf16), _ptr(h_f32), _ptr(mean), _ptr(rstd), c_float(eps), _stream()),
          "kr_rmsnorm_resid_ln_fwd")


def rmsnorm_resid_bwd(dout, y, gain, dy_bf16, dgain, drop: Optional[DropSpec] = None,
                      dcol: Optional[torch.Tensor] = None):
    """dcol [D] += column sums of dy (the bias gradient of the Linear that produced y)."""
    N, D = y.shape
    check(lib().kr_rmsnorm_resid_bwd(_ptr(dout), _ptr(y), _ptr(gain), _ptr(dy_bf16), _ptr(dgain), c_int(N),
                                     c_int(D), _dref(drop), _ptr(dcol), _stream()), "kr_rmsnorm_resid_bwd")


def _parts3(ts):
    ts = list(ts) + [None] * (3 - len(ts))
    return ts


def qkv_prep_fwd(ins, outs, gains, rope_mask: int, cos_t, sin_t, N: int, S: int, H: int):
    """ins/outs: lists (1..3) of [N, H*64] bf16 column-block views sharing one leading dimension."""
    n = len(ins)
    ld_in, ld_out = ins[0].stride(0), outs[0].stride(0)
    i3, o3, g3 = _parts3(ins), _parts3(outs), _parts3(gains)
    check(lib().kr_qkv_prep_fwd(_ptr(i3[0]), _ptr(i3[1]), _ptr(i3[2]), _ptr(o3[0]), _ptr(o3[1]), _ptr(o3[2]),
                                c_ll(ld_in), c_ll(ld_out), _ptr(g3[0]), _ptr(g3[1]), _ptr(g3[2]), c_int(n),
                                c_int(rope_mask), _ptr(cos_t), _ptr(sin_t), c_int(N), c_int(S), c_int(H),
                                _stream()), "kr_qkv_prep_fwd")


def qkv_prep_bwd(ins, grads, outs, gains, dgains, rope_mask: int, cos_t, sin_t, N: int, S: int, H: int):
    n = len(ins)
    ld_in, ld_out = ins[0].stride(0), outs[0].stride(0)
    f32_mask = 0
    for i, g in enumerate(grads):
        if g.dtype == torch.float32:
            f32_mask |= 1 << i
    i3, g3, o3, w3, d3 = _parts3(ins), _parts3(grads), _parts3(outs), _parts3(gains), _parts3(dgains)
    ldg = [g.stride(0) if g is not None else 0 for g in g3]
    check(lib().kr_qkv_prep_bwd(_ptr(i3[0]), _ptr(i3[1]), _ptr(i3[2]), _ptr(g3[0]), _ptr(g3[1]), _ptr(g3[2]),
                                c_ll(ldg[0]), c_ll(ldg[1]), c_ll(ldg[2]), _ptr(o3[0]), _ptr(o3[1]), _ptr(o3[2]),
                                c_ll(ld_in), c_ll(ld_out), _ptr(w3[0]), _ptr(w3[1]), _ptr(w3[2]), _ptr(d3[0]),
                                _ptr(d3[1]), _ptr(d3[2]), c_int(n), c_int(rope_mask), c_int(f32_mask),
                                _ptr(cos_t), _ptr(sin_t), c_int(N), c_int(S), c_int(H), _stream()),
          "kr_qkv_prep_bwd")


def glu_fwd(h, u, drop: Optional[DropSpec] = None):
    N, FF = u.shape
    check(lib().kr_glu_fwd(_ptr(h), _ptr(u), c_int(N), c_int(FF), _dref(drop), _stream()), "kr_glu_fwd")


def glu_bwd(du, h, dh, drop: Optional[DropSpec] = None):
    N, FF = du.shape
    check(lib().kr_glu_bwd(_ptr(du), _ptr(h), _ptr(dh), c_int(N), c_int(FF), _dref(drop), _stream()), "kr_glu_bwd")


def colsum_bf16(x, out):
    N, C = x.shape
    check(lib().kr_colsum_bf16(_ptr(x), c_ll(x.stride(0)), _ptr(out), c_int(N), c_int(C), _stream()),
          "kr_colsum_bf16")


def embed_fwd(idx, stress, emb, semb, pe, x, P: int, drop: Optional[DropSpec] = None):
    N, D = x.shape
    check(lib().kr_embed_fwd(_ptr(idx), _ptr(stress), _ptr(emb), _ptr(semb), _ptr(pe), _ptr(x), c_int(N),
                             c_int(P), c_int(D), _dref(drop), _stream()), "kr_embed_fwd")


def embed_bwd(dx, idx, stress, demb, dsemb, drop: Optional[DropSpec] = None):
    N, D = dx.shape
    check(lib().kr_embed_bwd(_ptr(dx), _ptr(idx), _ptr(stress), _ptr(demb), _ptr(dsemb), c_int(N), c_int(D),
                             _dref(drop), _stream()), "kr_embed_bwd")


def shift_cast(mel, out):
    B, T, C = mel.shape
    check(lib().kr_shift_cast(_ptr(mel), _ptr(out), c_int(B), c_int(T), c_int(C), _stream()), "kr_shift_cast")


def cast_bf16(src, dst):
    check(lib().kr_cast_bf16(_ptr(src), _ptr(dst), c_ll(src.numel()), _stream()), "kr_cast_bf16")


def scatter_rows(src, row_map, dst):
    R, C = src.shape
    check(lib().kr_scatter_rows(_ptr(src), _ptr(row_map), _ptr(dst), c_int(R), c_int(C),
                                c_int(int(dst.dtype == torch.bfloat16)), _stream()), "kr_scatter_rows")


def gather_rows(src, row_map, dst):
    R, C = dst.shape
    check(lib().kr_gather_rows(_ptr(src), _ptr(row_map), _ptr(dst), c_int(R), c_int(C), _stream()),
          "kr_gather_rows")


def eq_mask(idx, value: int, out):
    check(lib().kr_eq_mask_i64(_ptr(idx), c_ll(value), _ptr(out), c_ll(idx.numel()), _stream()), "kr_eq_mask_i64")


def nonfinite_flag(x, flag, bit: int):
    check(lib().kr_nonfinite_flag(_ptr(x), c_ll(x.numel()), _ptr(flag), c_int(bit), _stream()), "kr_nonfinite_flag")


def lr_index(dur, idx, lengths):
    B, P = dur.shape
    check(lib().kr_lr_index(_ptr(dur), _ptr(idx), _ptr(lengths), c_int(B), c_int(P), c_int(idx.shape[1]),
                            _stream()), "kr_lr_index")


def lr_index_masked(dur, pad_mask, idx, lengths):
    """`length_regulate` fallback indices (reference utils/lengths.py:108-153)."""
    B, P = dur.shape
    assert pad_mask.dtype == torch.uint8 and pad_mask.shape == dur.shape
    check(lib().kr_lr_index_masked(_ptr(dur), _ptr(pad_mask), _ptr(idx), _ptr(lengths), c_int(B), c_int(P),
                                   c_int(idx.shape[1]), _stream()), "kr_lr_index_masked")


def expand_rows_fwd(x, idx, out, frame_mask):
    B, P, D = x.shape
    check(lib().kr_expand_rows_fwd(_ptr(x), _ptr(idx), _ptr(out), _ptr(frame_mask), c_int(B), c_int(P),
                                   c_int(idx.shape[1]), c_int(D), _stream()), "kr_expand_rows_fwd")


def expand_rows_bwd(dout, idx, lengths, dx):
    B, P, D = dx.shape
    check(lib().kr_expand_rows_bwd(_ptr(dout), _ptr(idx), _ptr(lengths), _ptr(dx), c_int(B), c_int(P),
                                   c_int(idx.shape[1]), c_int(D), _stream()), "kr_expand_rows_bwd")


def range_flag(x, flag):
    check(lib().kr_range_flag(_ptr(x), c_ll(x.numel()), _ptr(flag), _stream()), "kr_range_flag")


def expand_adapt(enc, idx, lengths, pitch, energy, flags, pbins, ebins, pemb, eemb, row_of_tok, xpad, mem,
                 p_idx, e_idx, fmask_t, fmask_p, B, P, D, Tp, T):
    check(lib().kr_expand_adapt(_ptr(enc), _ptr(idx), _ptr(lengths), _ptr(pitch), _ptr(energy), _ptr(flags),
                                _ptr(pbins), _ptr(ebins), _ptr(pemb), _ptr(eemb), _ptr(row_of_tok), _ptr(xpad),
                                _ptr(mem), _ptr(p_idx), _ptr(e_idx), _ptr(fmask_t), _ptr(fmask_p), c_int(B),
                                c_int(P), c_int(D), c_int(Tp), c_int(T), c_int(pitch.shape[1]),
                                c_int(pbins.numel()), _stream()), "kr_expand_adapt")


def adapt_bwd(dmem, p_idx, e_idx, dpemb, deemb):
    rows, D = dmem.shape
    check(lib().kr_adapt_bwd(_ptr(dmem), _ptr(p_idx), _ptr(e_idx), _ptr(dpemb), _ptr(deemb), c_ll(rows), c_int(D),
                             _stream()), "kr_adapt_bwd")


def gn_fwd(x, row_group, group_rows, stats, gamma, beta, out, drop: Optional[DropSpec] = None):
    R, C = x.shape
    check(lib().kr_gn_fwd(_ptr(x), _ptr(row_group), _ptr(group_rows), _ptr(stats), _ptr(gamma), _ptr(beta),
                          _ptr(out), c_int(R), c_int(C), c_int(group_rows.numel()), _dref(drop), _stream()),
          "kr_gn_fwd")


def gn_bwd(dy, x, row_group, group_rows, stats, gsum, gamma, beta, dx, dgamma, dbeta,
           drop: Optional[DropSpec] = None):
    R, C = x.shape
    check(lib().kr_gn_bwd(_ptr(dy), _ptr(x), _ptr(row_group), _ptr(group_rows), _ptr(stats), _ptr(gsum),
                          _ptr(gamma), _ptr(beta), _ptr(dx), _ptr(dgamma), _ptr(dbeta), c_int(R), c_int(C),
                          c_int(group_rows.numel()), _dref(drop), _stream()), "kr_gn_bwd")


def drop_begin(state, path_site, path_p, table, B: int):
    """New dropout step: state[1] += 1, table[s, b] = stochastic-depth factor of branch s / sample b."""
    n = path_site.numel()
    check(lib().kr_drop_begin(_ptr(state), _ptr(path_site), _ptr(path_p), _ptr(table), c_int(n), c_int(B), _stream()),
          "kr_drop_begin")


def dec_in_drop(t, pe, y, T: int, drop: DropSpec, scale_a: float):
    N, D = t.shape
    check(lib().kr_dec_in_drop(_ptr(t), _ptr(pe), _ptr(y), c_int(N), c_int(T), c_int(D), c_float(scale_a),
                               _dref(drop), _stream()), "kr_dec_in_drop")


def drop_export_mask(state, site: int, p: float, rows: int, cols: int, ld: int, out, byte_lanes: bool = False):
    """out[rows, cols] uint8 = keep mask of `site` for elements r*ld + c (test aid)."""
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == rows * cols
    thr = drop_thr8(p) if byte_lanes else drop_thr(p)
    check(lib().kr_drop_export_mask(_ptr(state), ctypes.c_uint(site), ctypes.c_uint(thr), c_ll(rows),
                                    c_int(cols), c_ll(ld), _ptr(out), c_int(int(byte_lanes)), _stream()),
          "kr_drop_export_mask")


def vp_head_fwd(h, row_of_tok, w, b, mask, out, L: int, chunk: int):
    F = h.shape[1]
    check(lib().kr_vp_head_fwd(_ptr(h), _ptr(row_of_tok), _ptr(w), _ptr(b), _ptr(mask), _ptr(out),
                               c_int(out.numel()), c_int(L), c_int(F), c_int(chunk), _stream()), "kr_vp_head_fwd")


def vp_head_bwd(dout, h, tok_of_row, w, mask, dh, dw, db, L: int, chunk: int):
    R, F = h.shape
    check(lib().kr_vp_head_bwd(_ptr(dout), _ptr(h), _ptr(tok_of_row), _ptr(w), _ptr(mask), _ptr(dh), _ptr(dw),
                               _ptr(db), c_int(R), c_int(L), c_int(F), c_int(chunk), _stream()), "kr_vp_head_bwd")


def conv_dgrad_shadow(w2, wd, Co: int, Ci: int):
    check(lib().kr_conv_dgrad_shadow(_ptr(w2), _ptr(wd), c_int(Co), c_int(Ci), _stream()), "kr_conv_dgrad_shadow")


def conv_dgrad_shadow_multi(pairs, Co: int, Ci: int):
    """pairs: up to 8 (master conv weight fp32, dgrad shadow bf16) tensors of one (Co, Ci) shape, refreshed in one launch."""
    n = len(pairs)
    w2 = (c_void_p * n)(*[p[0].data_ptr() for p in pairs])
    wd = (c_void_p * n)(*[p[1].data_ptr() for p in pairs])
    check(lib().kr_conv_dgrad_shadow_multi(w2, wd, c_int(n), c_int(Co), c_int(Ci), _stream()), "kr_conv_dgrad_shadow_multi")


def stop_head_fwd(x, w, b, z):
    N, D = x.shape
    check(lib().kr_stop_head_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(z), c_int(N), c_int(D), _stream()),
          "kr_stop_head_fwd")


def stop_head_bwd(dz, x, dw, db):
    N, D = x.shape
    check(lib().kr_stop_head_bwd(_ptr(dz), _ptr(x), _ptr(dw), _ptr(db), c_int(N), c_int(D), _stream()),
          "kr_stop_head_bwd")


def losses_fwd_bwd(mel_pred, mel_tgt, dur_pred, dur_tgt, stop_pred, stop_tgt, pitch_pred, pitch_tgt,
                   energy_pred, energy_tgt, mel_len, ph_len, weights, pos_weight, delta_var, loss_scale, acc,
                   losses, dmel, ddur, dstop, dpitch, denergy):
    B, T, C = mel_tgt.shape
    P = dur_tgt.shape[1]
    Tp = pitch_pred.shape[1] if pitch_pred is not None else 0
    Tt = pitch_tgt.shape[1] if pitch_tgt is not None else 0
    w_dur, w_stop, w_pitch, w_energy = weights
    check(lib().kr_losses_fwd_bwd(_ptr(mel_pred), _ptr(mel_tgt), _ptr(dur_pred), _ptr(dur_tgt), _ptr(stop_pred),
                                  _ptr(stop_tgt), _ptr(pitch_pred), _ptr(pitch_tgt), _ptr(energy_pred),
                                  _ptr(energy_tgt), _ptr(mel_len), _ptr(ph_len), c_int(B), c_int(T), c_int(P),
                                  c_int(C), c_int(Tp), c_int(Tt), c_float(w_dur), c_float(w_stop),
                                  c_float(w_pitch), c_float(w_energy), c_float(pos_weight), c_float(delta_var),
                                  _ptr(loss_scale), _ptr(acc), _ptr(losses), _ptr(dmel), _ptr(ddur), _ptr(dstop),
                                  _ptr(dpitch), _ptr(denergy), _stream()), "kr_losses_fwd_bwd")


def spec_augment(x, spans, n_time: int, n_feat: int):
    """x: [B, T, D] bf16 or fp32 (in place); spans: int32 [B, n_time + n_feat, 2]."""
    B, T, D = x.shape
    check(lib().kr_spec_augment(_ptr(x), c_int(int(x.dtype == torch.float32)), _ptr(spans), c_int(B), c_int(T),
                                c_int(D), c_int(n_time), c_int(n_feat), _stream()), "kr_spec_augment")


def average_by_duration(values, dur, mask, label, out):
    B, P = dur.shape
    T = values.shape[1]
    check(lib().kr_average_by_duration(_ptr(values), _ptr(dur), _ptr(mask), _ptr(label), _ptr(out), c_int(B), c_int(P),
                                       c_int(T), _stream()), "kr_average_by_duration")


def val_metrics_acc_floats() -> int:
    return int(lib().kr_val_metrics_acc_floats())


def val_metrics(mel_pred, mel_tgt, pitch_pred, pitch_tgt, mel_lengths, acc):
    """Folds one validation batch into acc (see kr_val_metrics)."""
    B, T, C = mel_tgt.shape
    mel_pred, mel_tgt = mel_pred.contiguous(), mel_tgt.contiguous()
    Tp = 0
    if pitch_pred is not None:
        pitch_pred = pitch_pred.contiguous()
        Tp = pitch_pred.shape[1]
    if pitch_tgt is not None:
        pitch_tgt = pitch_tgt.contiguous()
    assert mel_lengths.dtype == torch.int64 and acc.dtype == torch.float32
    check(lib().kr_val_metrics(_ptr(mel_pred), _ptr(mel_tgt), _ptr(pitch_pred), _ptr(pitch_tgt), _ptr(mel_lengths.contiguous()),
                               _ptr(acc), c_int(B), c_int(T), c_int(Tp), c_int(C), _stream()), "kr_val_metrics")


def zero_(t: torch.Tensor) -> torch.Tensor:
    """In-place zero fill of a contiguous tensor through cudaMemsetAsync (no fill kernel)."""
    assert t.is_contiguous()
    check(lib().kr_memset_zero(_ptr(t), c_ll(t.numel() * t.element_size()), _stream()), "kr_memset_zero")
    return t
