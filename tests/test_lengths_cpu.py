"""Oracle restatements of the reference length-regulation functions vs fixtures generated from the LIVE
reference (tests/golden/make_golden_lengths.py): exact equality (gathers of fp32 values)."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _fix():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(HERE, "golden", "lengths.npz")).items()}


def test_expand_tokens_matches_reference_outputs():
    from oracle import acoustic as oa
    f = _fix()
    assert torch.equal(oa.expand_tokens(f["enc"], f["dur"]), f["exp_a"])
    assert torch.equal(oa.expand_tokens(f["enc"], f["dur"], 120), f["exp_b"])
    assert torch.equal(oa.expand_tokens(f["enc"][..., 0], f["dur"]), f["exp_c"])


def test_length_regulate_fallback_matches_reference_outputs():
    from oracle import acoustic as oa
    f = _fix()
    out, mask = oa.length_regulate_fallback(f["enc"], f["dur"].float(), f["pad"])
    assert out.shape == f["fb_out"].shape
    assert torch.equal(out, f["fb_out"]) and torch.equal(mask, f["fb_mask"])


# ---- edge cases side by side with the INSTALLED reference (baseline/_ref) ------------------------------------------------
def _ref_lengths():
    import logging
    import sys
    import pytest
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.utils import lengths
    return lengths


def _edge_durations():
    g = torch.Generator().manual_seed(9)
    yield "all zero", torch.zeros(2, 5, dtype=torch.long), None
    yield "one row zero", torch.tensor([[0, 0, 0, 0], [3, 0, 2, 1]]), None
    yield "single token", torch.tensor([[7]]), None
    yield "negative durations", torch.tensor([[2, -3, 4], [-1, -1, 5]]), None
    yield "max_len below the longest", torch.randint(0, 9, (3, 11), generator=g), 17
    yield "max_len above the longest", torch.randint(0, 9, (3, 11), generator=g), 200
    yield "max_len 1", torch.randint(1, 4, (2, 6), generator=g), 1
    yield "long utterance (config 4)", torch.randint(1, 16, (2, 260), generator=g), 2000
    yield "one huge duration", torch.tensor([[1, 1500, 1], [2, 2, 2]]), None


def test_expand_tokens_edge_cases_equal_the_installed_reference():
    """vectorized_expand_tokens (utils/lengths.py:16-96) on zero / negative / single / clipped / 2000-frame durations: the
    oracle's restatement returns the very same tensors (bit-exact: gathers of fp32 values), 3-D and 2-D tokens."""
    from oracle import acoustic as oa
    ref = _ref_lengths()
    g = torch.Generator().manual_seed(1)
    for label, dur, max_len in _edge_durations():
        B, P = dur.shape
        tok = torch.randn(B, P, 8, generator=g)
        for t in (tok, tok[..., 0]):
            want = ref.vectorized_expand_tokens(t, dur, max_len=max_len)
            got = oa.expand_tokens(t, dur, max_len)
            assert got.shape == want.shape and torch.equal(got, want), (label, t.dim(), got.shape, want.shape)
