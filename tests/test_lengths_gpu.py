"""kokoro_ruslan_b200.lengths (CUDA) vs the reference fixtures and the oracle: indices and gathers are exact,
the fallback's backward (segment sums) matches autograd of the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _fix():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(HERE, "golden", "lengths.npz")).items()}


def test_vectorized_expand_and_module_match_reference():
    from kokoro_ruslan_b200.lengths import LengthRegulator, vectorized_expand_tokens
    f = _fix()
    enc, dur = f["enc"].cuda(), f["dur"].cuda()
    assert torch.equal(vectorized_expand_tokens(enc, dur).cpu(), f["exp_a"])
    assert torch.equal(LengthRegulator()(enc, dur, max_len=120).cpu(), f["exp_b"])
    assert torch.equal(vectorized_expand_tokens(enc[..., 0].contiguous(), dur).cpu(), f["exp_c"])
    x = enc.clone().requires_grad_(True)
    assert not vectorized_expand_tokens(x, dur).requires_grad          # the reference detaches (lengths.py:30)


def test_length_regulate_fallback_forward_backward():
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.lengths import length_regulate
    f = _fix()
    x = f["enc"].cuda().requires_grad_(True)
    out, mask = length_regulate(x, f["dur"].float().cuda(), f["pad"].cuda())
    assert torch.equal(out.detach().cpu(), f["fb_out"]) and torch.equal(mask.cpu(), f["fb_mask"])
    g = torch.Generator().manual_seed(3)
    dout = torch.randn(out.shape, generator=g)
    out.backward(dout.cuda())
    xo = f["enc"].clone().requires_grad_(True)
    oo, _ = oa.length_regulate_fallback(xo, f["dur"].float(), f["pad"])
    oo.backward(dout)
    assert torch.allclose(x.grad.cpu(), xo.grad, rtol=1e-6, atol=1e-6)
    assert float(x.grad[3].abs().max()) == 0.0                           # fully padded sample gets no gradient


def test_long_utterances_bit_exact_indices():
    """BASELINE config 4: mel_frames ~ 2000, min_batch 1 — index tensor bit-exact against the oracle."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200 import ops
    g = torch.Generator().manual_seed(9)
    dur = torch.randint(0, 40, (1, 120), generator=g)
    pad = torch.zeros(1, 120, dtype=torch.bool)
    pad[0, 100:] = True
    d_eff = torch.where(pad, torch.zeros_like(dur), dur.clamp(min=1))
    want, L = oa.length_regulate_index(d_eff)
    idx = torch.empty(1, want.shape[1], dtype=torch.int32, device="cuda")
    lens = torch.empty(1, dtype=torch.int32, device="cuda")
    ops.lr_index_masked(dur.cuda(), pad.to(torch.uint8).cuda(), idx, lens)
    assert torch.equal(idx.cpu().long(), want) and int(lens[0]) == int(L[0]) and int(L[0]) > 1500

