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
