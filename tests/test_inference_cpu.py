"""Oracle restatement of the reference's autoregressive inference (KV-cached frame-by-frame decode, predicted
durations / pitch / energy, stop rules) against fixtures from the LIVE reference
(tests/golden/make_golden_inference.py): same number of generated frames, same mel within fp32 noise."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _setup():
    from oracle import acoustic as oa
    f = np.load(os.path.join(HERE, "golden", "inference.npz"))
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                            variance_filter=64, max_len=1200)
    sd = oa.seeded_state_dict(cfg, seed=int(f["seed"]))
    sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([float(f["dur_bias"])])
    sd["stop_token_predictor.bias"] = torch.tensor([float(f["stop_bias"])])
    return f, cfg, sd


def test_single_utterance_generation_matches_reference():
    from oracle import inference as oi
    f, cfg, sd = _setup()
    mel, probs = oi.forward_inference(sd, cfg, torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"]),
                                      return_stop_probs=True)
    want = torch.from_numpy(f["mel"])
    assert mel.shape == want.shape and mel.shape[1] >= 12
    assert float((mel - want).abs().max()) < 1e-4
    assert len(probs) == mel.shape[1] and all(0.0 < p < 1.0 for p in probs)


def test_padded_batch_generation_matches_reference():
    from oracle import inference as oi
    f, cfg, sd = _setup()
    mel = oi.forward_inference(sd, cfg, torch.from_numpy(f["idx2"]), None, stop_threshold=0.45)
    want = torch.from_numpy(f["mel2"])
    assert mel.shape == want.shape
    assert float((mel - want).abs().max()) < 1e-4


def test_generation_bounds_follow_the_reference_formula():
    from oracle.inference import generation_bounds
    assert generation_bounds(100) == (70, 300)
    assert generation_bounds(5) == (12, 85)
    assert generation_bounds(1000) == (700, 1600)
    assert generation_bounds(3000, max_len=2000) == (2100, 2000 if 2000 > 2100 else min(2000, 2101))


def test_rope_quirk_query_is_always_position_zero():
    """The cached-decode query is rotated as position 0 (reference transformers.py:276-277): documenting the
    train / inference mismatch the CUDA decode kernel will have to reproduce."""
    from oracle import inference as oi
    t = torch.randn(1, 2, 1, 64)
    assert torch.allclose(oi._rope_at(t, 0), t, atol=1e-6)            # position 0 = identity rotation
    assert not torch.allclose(oi._rope_at(t, 7), t, atol=1e-3)
