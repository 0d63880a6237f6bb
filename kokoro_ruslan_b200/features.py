"""On-device log-mel feature extraction — host-side mirror of the reference's per-item feature code
(src/kokoro/data/dataset.py:162-178 ``torchaudio.transforms.MelSpectrogram`` construction, :672 peak
normalisation, :687-697 short-clip padding, transform and ``log(x + 1e-9)``).

The reference runs this on the CPU inside ``Dataset.__getitem__`` one utterance at a time; here a whole
batch of padded waveforms is processed by one launch of ``kr_mel_stft`` (one CTA per frame).
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional

import torch

from ._lib import check, lib
from .ops import _ptr, _stream


def mel_filterbank_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """[n_freqs, n_mels] HTK triangles without area normalisation (what
    ``MelSpectrogram(mel_scale='htk', norm=None)`` builds)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs, dtype=torch.float64)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2, dtype=torch.float64)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.minimum(down, up), min=0.0).to(torch.float32)


class LogMelSpectrogram:
    """``LogMelSpectrogram(config)(wav, lengths)`` -> (B, n_mels, 1 + N // hop) log-mel on the device."""

    def __init__(self, sample_rate: int = 22050, n_fft: int = 1024, win_length: int = 1024, hop_length: int = 256,
                 n_mels: int = 80, f_min: float = 0.0, f_max: float = 8000.0, device="cuda", log_eps: float = 1e-9):
        if not torch.cuda.is_available():
            raise RuntimeError("LogMelSpectrogram needs a CUDA device (no CPU fallback)")
        if n_fft <= 0 or hop_length <= 0:
            raise ValueError("n_fft and hop_length must be positive integers.")     # dataset.py:150-158
        if (n_fft, win_length, hop_length) != (1024, 1024, 256):
            raise RuntimeError("the sm_100a mel-STFT kernel is built for n_fft = win = 1024, hop = 256")
        self.n_fft, self.hop, self.n_mels, self.log_eps = n_fft, hop_length, n_mels, log_eps
        self.device = torch.device(device)
        fb = mel_filterbank_htk(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate)
        self.fb_t = fb.t().contiguous().to(self.device)            # [n_mels, 513]

    def __call__(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 peak_normalize: bool = True) -> torch.Tensor:
        if wav.dim() == 1:
            wav = wav.unsqueeze(0)
        wav = wav.to(self.device, torch.float32)
        B, n_max = wav.shape
        if n_max < self.n_fft:                                      # dataset.py:687-690
            wav = torch.nn.functional.pad(wav, (0, self.n_fft - n_max))
            if lengths is not None:
                lengths = torch.clamp(lengths, min=self.n_fft)
            n_max = self.n_fft
        wav = wav.contiguous()
        if lengths is not None:
            lengths = lengths.to(self.device, torch.int64).contiguous()
        frames = 1 + n_max // self.hop
        out = torch.empty(B, self.n_mels, frames, dtype=torch.float32, device=self.device)
        peak = None
        if peak_normalize:
            peak = torch.empty(B, dtype=torch.float32, device=self.device)
            check(lib().kr_wave_peak(_ptr(wav), _ptr(lengths), _ptr(peak), ctypes.c_int(B), ctypes.c_longlong(n_max),
                                     _stream()), "kr_wave_peak")
        check(lib().kr_mel_stft(_ptr(wav), _ptr(lengths), _ptr(peak), _ptr(self.fb_t), _ptr(out), ctypes.c_int(B),
                                ctypes.c_longlong(n_max), ctypes.c_int(frames), ctypes.c_int(self.n_mels),
                                ctypes.c_int(self.n_fft), ctypes.c_int(self.hop), ctypes.c_float(self.log_eps),
                                _stream()), "kr_mel_stft")
        return out
