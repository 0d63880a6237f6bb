"""Oracle restatements of the reference's pitch (YIN / CMND) and energy extractors — the next row of the feature
pipeline — against fixtures generated from the LIVE reference
(tests/golden/make_golden_features.py).  These are chains of thresholded decisions, hence the two-level gate:
identical on >= 99 % of the frames, and never far off."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _fix():
    return np.load(os.path.join(HERE, "golden", "features.npz"))


def test_pitch_extractor_matches_reference():
    from oracle import features as of
    f = _fix()
    got = of.extract_pitch(f["wav"])
    assert got.shape == f["pitch"].shape
    d = np.abs(got - f["pitch"])
    assert (d < 1e-4).mean() >= 0.99 and d.max() < 0.05, ((d < 1e-4).mean(), d.max())
    assert ((got > 0) == (f["pitch"] > 0)).mean() >= 0.99                      # voicing decisions
    assert 0.3 < (f["pitch"] > 0).mean() < 0.95                                 # the fixture has voiced AND unvoiced frames
    short = of.extract_pitch(f["wav"][0, :1500])                                # shorter than one analysis window
    assert short.shape == f["pitch_short"].shape and np.abs(short - f["pitch_short"]).max() < 1e-4


def test_energy_extractor_matches_reference():
    from oracle import features as of
    f = _fix()
    assert np.abs(of.extract_energy_from_mel(f["mel"]) - f["e_log"]).max() < 1e-5            # log-mel (heuristic branch)
    assert np.abs(of.extract_energy_from_mel(np.exp(f["mel"]), False) - f["e_lin"]).max() < 1e-5
    assert np.abs(of.extract_energy_from_mel(f["mel"][:, :2]) - f["e_short"]).max() < 1e-6   # < 3 frames: min / max


def test_extractors_equal_the_installed_reference_on_edge_signals():
    """Side by side with the installed reference's PitchExtractor / EnergyExtractor (model/variance_predictor.py:448-688) on
    signals the fixture does not hold: digital silence, a pure tone, a tone with a DC offset, white noise, a clipped square
    wave, a 300-sample stub; energy of a constant mel and of a single frame."""
    import logging
    import sys
    import pytest
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import features as of
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.model.variance_predictor import EnergyExtractor, PitchExtractor
    sr, n = 22050, 22050
    t = np.arange(n, dtype=np.float32) / sr
    rng = np.random.default_rng(5)
    signals = {
        "silence": np.zeros(n, np.float32),
        "220 Hz tone": (0.4 * np.sin(2 * np.pi * 220.0 * t)).astype(np.float32),
        "tone + DC": (0.3 * np.sin(2 * np.pi * 130.0 * t) + 0.2).astype(np.float32),
        "white noise": (0.1 * rng.standard_normal(n)).astype(np.float32),
        "clipped square": np.clip(4.0 * np.sin(2 * np.pi * 95.0 * t), -1.0, 1.0).astype(np.float32),
        "300-sample stub": (0.2 * np.sin(2 * np.pi * 300.0 * t[:300])).astype(np.float32),
    }
    for label, w in signals.items():
        want = PitchExtractor.extract_pitch(torch.from_numpy(w)).numpy()
        got = of.extract_pitch(w)
        got = got.reshape(want.shape) if got.size == want.size else got
        assert got.shape == want.shape, (label, got.shape, want.shape)
        d = np.abs(got - want)
        assert (d < 1e-4).mean() >= 0.99 and d.max() < 0.05, (label, float((d < 1e-4).mean()), float(d.max()))
    mels = {"constant mel": np.full((1, 80, 40), -4.0, np.float32), "single frame": rng.standard_normal((1, 80, 1)).astype(np.float32),
            "two utterances": (rng.standard_normal((2, 80, 33)) * 2.0 - 5.0).astype(np.float32)}
    for label, m in mels.items():
        want = EnergyExtractor.extract_energy_from_mel(torch.from_numpy(m)).numpy()
        got = of.extract_energy_from_mel(m)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-5, (label, float(np.abs(got - want).max()))
