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
