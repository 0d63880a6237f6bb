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


def test_oracle_equals_torchaudio_on_edge_waveforms():
    """Side by side with torchaudio's own MelSpectrogram (configured as the reference's data/dataset.py:162-178, with the
    reference's peak normalisation and short-input padding :687-690) on waveforms the fixture does not hold: digital silence,
    a single click, a waveform shorter than one FFT window, exactly one window, a full-scale square wave."""
    import pytest
    torchaudio = pytest.importorskip("torchaudio")
    from oracle import melstft as om
    tr = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, n_mels=80, hop_length=256, win_length=1024,
                                              f_min=0.0, f_max=8000.0, power=2.0, normalized=False, window_fn=torch.hann_window)
    t = torch.arange(6000) / 22050.0
    click = torch.zeros(4000)
    click[1777] = 1.0
    waves = {"silence": torch.zeros(5000), "click": click, "shorter than a window": 0.3 * torch.sin(2 * torch.pi * 440.0 * t[:700]),
             "exactly one window": 0.3 * torch.sin(2 * torch.pi * 440.0 * t[:1024]),
             "full-scale square": torch.sign(torch.sin(2 * torch.pi * 150.0 * t))}
    for label, x in waves.items():
        xn = x / (x.abs().max() + 1e-9)
        if xn.numel() < 1024:
            xn = torch.nn.functional.pad(xn, (0, 1024 - xn.numel()))
        want = torch.log(tr(xn.unsqueeze(0)).squeeze(0) + 1e-9).numpy()
        got = om.log_mel(x.numpy())
        assert got.shape == want.shape, (label, got.shape, want.shape)
        # torchaudio computes in fp32: bins more than ~4 decades below the loudest one (leakage valleys of a pure tone, the
        # 1e-9 floor) carry its rounding noise, so the log-domain gate covers the bins above that and the linear gate all
        loud = want > want.max() - 9.0
        assert np.abs(got - want)[loud].max(initial=0.0) < 2e-4, (label, float(np.abs(got - want)[loud].max(initial=0.0)))
        assert np.abs(np.exp(got) - np.exp(want)).max() < 1e-5 * max(1.0, float(np.exp(want).max())), label   # fp32 FFT
