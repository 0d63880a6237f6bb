"""Length regulation — drop-in for the reference's ``kokoro.utils.lengths`` entry points
(src/kokoro/utils/lengths.py): ``vectorized_expand_tokens`` / ``LengthRegulator`` (:16-105, the detached,
CPU-round-trip expansion used with the variance adaptor) and ``length_regulate`` (:108-153, the
fallback of ``use_variance_predictor=False``: padded tokens skipped, durations clamped to >= 1, autograd kept).

Index tensors are bit-exact restatements computed on the device (`kr_lr_index`, `kr_lr_index_masked`), the
expansion is an exact fp32 gather (`kr_expand_rows_fwd`) and the fallback's backward is a deterministic
segment sum (`kr_expand_rows_bwd`).  The one host synchronisation — the expanded length max_b sum(d) — is the
same one the reference performs (lengths.py:44-47, :127).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import ops


def _index(durations: torch.Tensor, pad_mask: Optional[torch.Tensor], max_len: Optional[int]):
    B, P = durations.shape
    dev = durations.device
    dur = durations.long().contiguous()
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    pm = None if pad_mask is None else pad_mask.to(torch.uint8).contiguous()
    if max_len is None:
        probe = torch.empty(B, 0, dtype=torch.int32, device=dev)
        (ops.lr_index if pm is None else ops.lr_index_masked)(*((dur, probe, lengths) if pm is None else (dur, pm, probe, lengths)))
        max_len = max(1, int(lengths.max().item()))
    idx = torch.empty(B, int(max_len), dtype=torch.int32, device=dev)
    if pm is None:
        ops.lr_index(dur, idx, lengths)
    else:
        ops.lr_index_masked(dur, pm, idx, lengths)
    return idx, lengths


def vectorized_expand_tokens(tokens: torch.Tensor, durations: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    """(B,P,D) or (B,P) tokens -> left-packed (B, T', D) / (B, T') expansion; gradients are cut like the reference
    (``tokens.detach()``, lengths.py:30)."""
    if not tokens.is_cuda:
        raise RuntimeError("kokoro_ruslan_b200.lengths needs CUDA tensors (no CPU fallback)")
    squeeze = tokens.dim() == 2
    x = tokens.detach().float()
    if squeeze:
        x = x.unsqueeze(-1).expand(-1, -1, 4)          # the gather moves float4 rows
    x = x.contiguous()
    idx, _ = _index(durations, None, max_len)
    out = torch.empty(x.shape[0], idx.shape[1], x.shape[2], dtype=torch.float32, device=x.device)
    ops.expand_rows_fwd(x, idx, out, None)
    out = out[..., 0] if squeeze else out
    return out.to(tokens.dtype)


class LengthRegulator:
    """``LengthRegulator().forward(x, durations, max_len)`` of the reference (lengths.py:99-105)."""

    def forward(self, x: torch.Tensor, durations: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
        return vectorized_expand_tokens(x, durations, max_len=max_len)

    __call__ = forward


class _LengthRegulateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, durations, pad_mask):
        x = enc.float().contiguous()
        idx, lengths = _index(durations, pad_mask, None)
        B, P, D = x.shape
        out = torch.empty(B, idx.shape[1], D, dtype=torch.float32, device=x.device)
        mask = torch.empty(B, idx.shape[1], dtype=torch.uint8, device=x.device)
        ops.expand_rows_fwd(x, idx, out, mask)
        ctx.save_for_backward(idx, lengths)
        ctx.shape = (B, P, D)
        ctx.dtype = enc.dtype
        ctx.mark_non_differentiable(mask)
        return out.to(enc.dtype), mask

    @staticmethod
    def backward(ctx, dout, _dmask):
        idx, lengths = ctx.saved_tensors
        dx = torch.empty(ctx.shape, dtype=torch.float32, device=dout.device)
        ops.expand_rows_bwd(dout.float().contiguous(), idx, lengths, dx)
        return dx.to(ctx.dtype), None, None


def length_regulate(encoder_outputs: torch.Tensor, durations: torch.Tensor,
                    text_padding_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Reference ``length_regulate`` (lengths.py:108-153): returns (expanded (B, T', D), frame_mask (B, T') bool,
    True = padding).  Differentiable w.r.t. ``encoder_outputs``."""
    if not encoder_outputs.is_cuda:
        raise RuntimeError("kokoro_ruslan_b200.lengths needs CUDA tensors (no CPU fallback)")
    out, mask = _LengthRegulateFn.apply(encoder_outputs, durations, text_padding_mask)
    return out, mask.bool()


def average_by_duration(values: torch.Tensor, durations: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Reference ``average_by_duration`` (lengths.py:156-208): frame-level (B, T) values -> token-level (B, P) means through
    the durations.  Bit-identical to the reference's scatter formulation, quirks included (frames no token covers are
    averaged into token 0; tokens starting past the last frame pile their labels onto frame T - 1) — see
    csrc/kr_lengths_core.cuh."""
    if not values.is_cuda:
        raise RuntimeError("kokoro_ruslan_b200.lengths needs CUDA tensors (no CPU fallback)")
    B, P = durations.shape
    T = values.shape[1]
    v = values.float().contiguous()
    dur = durations.long().contiguous()
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    label = torch.empty(B, T, dtype=torch.int32, device=v.device)
    out = torch.empty(B, P, dtype=torch.float32, device=v.device)
    ops.average_by_duration(v, dur, m, label, out)
    return out.to(values.dtype)
