"""ORACLE (test infrastructure, never the product path) — CPU restatement of the log-mel feature
extractor of igorshmukler/kokoro-ruslan (reference src/kokoro/data/dataset.py:162-178 transform,
:672 peak normalisation, :694-697 log) in float64 numpy.

torchaudio.transforms.MelSpectrogram(sample_rate 22050, n_fft 1024, win 1024, hop 256, f_min 0,
f_max 8000, n_mels 80, power 2, hann (periodic), center=True, pad_mode='reflect', mel_scale 'htk',
norm=None).  Parity is PINNED: tests/golden/melstft.npz holds torchaudio's own output for seeded
waveforms (tests/golden/make_golden_melstft.py).
"""
from __future__ import annotations

import numpy as np

SR, N_FFT, HOP, N_MELS, F_MIN, F_MAX = 22050, 1024, 256, 80, 0.0, 8000.0


def hz_to_mel_htk(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel_to_hz_htk(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(n_freqs: int = N_FFT // 2 + 1, f_min: float = F_MIN, f_max: float = F_MAX,
                   n_mels: int = N_MELS, sample_rate: int = SR) -> np.ndarray:
    """[n_freqs, n_mels] triangular filters, HTK scale, no area normalisation
    (torchaudio.functional.melscale_fbanks semantics)."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = mel_to_hz_htk(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]                 # [n_freqs, n_mels + 2]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(0.0, np.minimum(down, up))


def log_mel(wav: np.ndarray, peak_normalize: bool = True, log_eps: float = 1e-9) -> np.ndarray:
    """wav: (N,) -> (80, 1 + N // 256) float64."""
    x = np.asarray(wav, dtype=np.float64)
    if peak_normalize:
        x = x / (np.abs(x).max() + 1e-9)
    if x.shape[0] < N_FFT:                                        # dataset.py:687-690
        x = np.pad(x, (0, N_FFT - x.shape[0]))
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    n_frames = 1 + x.shape[0] // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(n_frames)[:, None]
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)
    spec = np.fft.rfft(xp[idx] * win[None, :], axis=1)
    power = spec.real ** 2 + spec.imag ** 2                       # [frames, 513]
    mel = power @ mel_filterbank()
    return np.log(mel + log_eps).T
