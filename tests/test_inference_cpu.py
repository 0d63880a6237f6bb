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


def test_oracle_equals_the_installed_reference_on_generation_edge_cases():
    """Side by side with the installed reference's forward_inference (baseline/_ref) on cases the fixture does not hold: a
    single phoneme, a stop head that never fires (generation ends at the upper bound), one that fires at once (ends at the
    lower bound), a small max_len, a padded batch of three.  Same number of frames, same mel within fp32 noise."""
    import logging
    import sys
    import pytest
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import acoustic as oa
    from oracle import inference as oi
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.model.model import KokoroModel
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=2, ff_dim=128, variance_filter=64,
                            max_len=1200)
    g = torch.Generator().manual_seed(31)
    idx3 = torch.randint(1, cfg.vocab_size, (3, 10), generator=g)
    idx3[1, 4:] = 0
    idx3[2, 7:] = 0
    cases = [("single phoneme", 1.2, -1.0, torch.randint(1, cfg.vocab_size, (1, 1), generator=g), {}),
             ("stop never fires", 0.8, -30.0, torch.randint(1, cfg.vocab_size, (1, 6), generator=g), {}),
             ("stop fires at once", 0.8, 30.0, torch.randint(1, cfg.vocab_size, (1, 6), generator=g), {}),
             ("small max_len", 1.5, -30.0, torch.randint(1, cfg.vocab_size, (1, 12), generator=g), dict(max_len=20)),
             ("padded batch of three", 1.0, -0.5, idx3, dict(stop_threshold=0.4))]
    for label, dur_bias, stop_bias, idx, kw in cases:
        sd = oa.seeded_state_dict(cfg, seed=6)
        sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([dur_bias])
        sd["stop_token_predictor.bias"] = torch.tensor([stop_bias])
        m = KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim, n_encoder_layers=1, n_heads=2,
                        encoder_ff_dim=cfg.ff_dim, encoder_dropout=0.1, decoder_dropout=0.1, decoder_input_dropout=0.1,
                        n_decoder_layers=2, decoder_ff_dim=cfg.ff_dim, max_decoder_seq_len=cfg.max_len,
                        variance_filter_size=cfg.variance_filter, variance_dropout=0.1, n_variance_bins=cfg.n_bins,
                        pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=True, qk_norm=True,
                        ffn_output_norm=True)
        m.load_state_dict(sd, strict=True)
        m.eval()
        with torch.no_grad():
            want = m.forward_inference(idx, stress_indices=None, **kw)
        got = oi.forward_inference(sd, cfg, idx, None, **kw)
        assert got.shape == want.shape, (label, got.shape, want.shape)
        assert float((got - want).abs().max()) < 1e-4, (label, float((got - want).abs().max()))
