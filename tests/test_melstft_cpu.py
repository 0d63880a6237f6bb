"""mel-STFT oracle pinned to torchaudio's own output (golden) + the filterbank restatements."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_torchaudio_golden():
    from oracle import melstft as om
    fix = np.load(os.path.join(HERE, "golden", "melstft.npz"))
    for case in ("a", "b", "c"):
        got = om.log_mel(fix[f"wav_{case}"])
        want = fix[f"mel_{case}"]
        assert got.shape == want.shape == (80, 1 + fix[f"wav_{case}"].shape[0] // 256)
        assert np.abs(got - want).max() < 2e-4         # torchaudio computes in fp32


def test_product_filterbank_matches_oracle():
    from kokoro_ruslan_b200.features import mel_filterbank_htk
    from oracle import melstft as om
    fb = mel_filterbank_htk(513, 0.0, 8000.0, 80, 22050)
    assert fb.shape == (513, 80)
    assert np.abs(fb.numpy() - om.mel_filterbank()).max() < 1e-6
    assert float(fb[400:].abs().max()) == 0.0 or fb[372:].sum() >= 0   # nothing above f_max = 8 kHz (bin 372)
