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
         accumulate: bool = False, splits: int = 1) -> torch.Tensor:
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
    rc = lib().kr_gemm_bf16(_ptr(a), _ptr(b), _ptr(out), c_int(M), c_int(N), c_int(K), c_int(batch),
                            c_ll(a2.stride(0)), c_ll(b2.stride(0)), c_ll(o2.stride(0)),
                            c_ll(sa), c_ll(sb), c_ll(sc), c_int(int(a_mn_major)),
                            c_int(int(b_mn_major)), c_int(epi), _ptr(bias), _ptr(resid),
                            c_ll(ldr), c_ll(sr), c_int(resid_mod), c_float(alpha), c_int(splits),
                            _stream())
    check(rc, "kr_gemm_bf16")
    return out


def _heads_strides(t: torch.Tensor):
    """t is a [B, S, H, 64] bf16 view with unit inner stride and head stride 64."""
    assert t.dtype == torch.bfloat16 and t.dim() == 4 and t.shape[3] == 64
    assert t.stride(3) == 1 and t.stride(2) == 64, t.stride()
    return c_ll(t.stride(1)), c_ll(t.stride(0))


def attn_fwd(q, k, v, o, lse, key_mask: Optional[torch.Tensor], causal: bool, scale: float):
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
                           c_int(int(causal)), c_float(scale), _stream())
    check(rc, "kr_attn_fwd")


def attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, key_mask, causal: bool, scale: float):
    """Flash attention backward.  dq: fp32 [B,Sq,H,64] and must be ZEROED by the caller (atomic
    accumulation); dk, dv: bf16 [B,Sk,H,64]; delta: fp32 scratch [B,H,Sq]."""
    B, Sq, H, _ = q.shape
    Sk = k.shape[1]
    assert dq.dtype == torch.float32 and dq.stride(3) == 1 and dq.stride(2) == 64
    rc = lib().kr_attn_bwd(_ptr(q), *_heads_strides(q), _ptr(k), *_heads_strides(k), _ptr(v),
                           *_heads_strides(v), _ptr(o), *_heads_strides(o), _ptr(d_o),
                           *_heads_strides(d_o), _ptr(lse), _ptr(delta), _ptr(dq),
                           c_ll(dq.stride(1)), c_ll(dq.stride(0)), _ptr(dk), *_heads_strides(dk),
                           _ptr(dv), *_heads_strides(dv), _ptr(key_mask), c_int(B), c_int(H),
                           c_int(Sq), c_int(Sk), c_int(int(causal)), c_float(scale), _stream())
    check(rc, "kr_attn_bwd")
