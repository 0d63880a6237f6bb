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
        # the filters are triangles: [first non-zero bin, one past the last) per mel row — the kernel skips the exact zeros
        nz = fb.t() != 0
        first = torch.where(nz.any(dim=1), nz.float().argmax(dim=1), torch.zeros(n_mels, dtype=torch.long))
        last = torch.where(nz.any(dim=1), nz.shape[1] - nz.flip(1).float().argmax(dim=1), torch.zeros(n_mels, dtype=torch.long))
        self.fb_ranges = torch.stack([first, last], dim=1).to(torch.int32).contiguous().to(self.device)

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
        check(lib().kr_mel_stft(_ptr(wav), _ptr(lengths), _ptr(peak), _ptr(self.fb_t), _ptr(self.fb_ranges), _ptr(out), ctypes.c_int(B),
                                ctypes.c_longlong(n_max), ctypes.c_int(frames), ctypes.c_int(self.n_mels),
                                ctypes.c_int(self.n_fft), ctypes.c_int(self.hop), ctypes.c_float(self.log_eps),
                                _stream()), "kr_mel_stft")
        return out


# ----------------------------------------------------------------------------------------------------------------
# Pitch / energy extractors — same class / method names and argument meaning as the reference's
# src/kokoro/model/variance_predictor.py:442-688, batched over padded utterances with per-item lengths.
# ----------------------------------------------------------------------------------------------------------------
def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (the feature kernels have no CPU fallback)")


class PitchExtractor:
    """``PitchExtractor.extract_pitch(waveform, ...)`` -> normalised F0 contour in [0, 1], 0 = unvoiced
    (variance_predictor.py:448-625).  ``lengths`` (optional, new) gives per-utterance sample counts of a padded batch:
    every row is then analysed exactly as if it had been passed alone, and frames beyond its own
    ``1 + max(len, 2048) // hop`` are zero."""

    WIN = 2048

    @staticmethod
    def num_frames(n_samples: int, hop_length: int = 256) -> int:
        return max(int(n_samples), PitchExtractor.WIN) // hop_length + 1

    @staticmethod
    def extract_pitch(waveform: torch.Tensor, sample_rate: int = 22050, hop_length: int = 256, fmin: float = 50.0,
                      fmax: float = 800.0, win_length: Optional[int] = None,
                      lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        _need_cuda(waveform, "PitchExtractor.extract_pitch")
        hop = max(1, int(hop_length))
        win = int(win_length) if win_length is not None else max(2048, hop * 8)          # :491-492
        if hop != 256 or win != PitchExtractor.WIN:
            raise RuntimeError("the sm_100a pitch kernel is built for hop 256 / analysis window 2048")
        squeeze = waveform.dim() == 1
        wav = (waveform.unsqueeze(0) if squeeze else waveform).to(torch.float32).contiguous()
        B, n_max = wav.shape
        if lengths is not None:
            lengths = lengths.to(wav.device, torch.int64).contiguous()
        frames = PitchExtractor.num_frames(n_max, hop)
        buf = torch.empty(5, B, frames, dtype=torch.float32, device=wav.device)          # cand, acmax, energy, work, out
        check(lib().kr_pitch_frames(_ptr(wav), _ptr(lengths), _ptr(buf[0]), _ptr(buf[1]), _ptr(buf[2]), ctypes.c_int(B),
                                    ctypes.c_longlong(n_max), ctypes.c_int(frames), ctypes.c_int(int(sample_rate)),
                                    ctypes.c_int(hop), ctypes.c_float(fmin), ctypes.c_float(fmax), _stream()),
              "kr_pitch_frames")
        check(lib().kr_pitch_track(_ptr(buf[0]), _ptr(buf[1]), _ptr(buf[2]), _ptr(lengths), _ptr(buf[3]), _ptr(buf[4]),
                                   ctypes.c_int(B), ctypes.c_longlong(n_max), ctypes.c_int(frames), ctypes.c_float(fmin),
                                   ctypes.c_float(fmax), _stream()), "kr_pitch_track")
        out = buf[4]
        return out[0] if squeeze else out


class EnergyExtractor:
    """``EnergyExtractor.extract_energy_from_mel(mel_spec, log_domain)`` -> energy contour in [0, 1]
    (variance_predictor.py:633-688).  ``mel_spec`` is (B, frames, n_mels) or (frames, n_mels) like the reference's;
    additions: ``frames`` = per-utterance valid frame counts of a padded batch, ``channel_major`` for the mel-STFT
    kernel's (B, n_mels, frames) layout, ``exp_input`` to feed that kernel's LOG-mel while asking for the linear-power
    semantics the dataset uses (data/dataset.py:813 passes the pre-log mel with ``log_domain=False``)."""

    @staticmethod
    def extract_energy_from_mel(mel_spec: torch.Tensor, log_domain: Optional[bool] = None,
                                frames: Optional[torch.Tensor] = None, channel_major: bool = False,
                                exp_input: bool = False) -> torch.Tensor:
        _need_cuda(mel_spec, "EnergyExtractor.extract_energy_from_mel")
        squeeze = mel_spec.dim() == 2
        mel = (mel_spec.unsqueeze(0) if squeeze else mel_spec).to(torch.float32).contiguous()
        if channel_major:
            B, M, T = mel.shape
        else:
            B, T, M = mel.shape
        if log_domain is None:
            # :653-657 `median() < -1` (lower median) == more than (n - 1) // 2 values are below -1: a count, not a sort
            log_domain = int((mel < -1.0).sum().item()) > (mel.numel() - 1) // 2
        if frames is not None:
            frames = frames.to(mel.device, torch.int64).contiguous()
        buf = torch.empty(2, B, T, dtype=torch.float32, device=mel.device)
        check(lib().kr_energy_frames(_ptr(mel), _ptr(buf[0]), ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(M),
                                     ctypes.c_int(0 if channel_major else 1), ctypes.c_int(int(exp_input)),
                                     ctypes.c_int(int(bool(log_domain))), _stream()), "kr_energy_frames")
        check(lib().kr_energy_norm(_ptr(buf[0]), _ptr(frames), _ptr(buf[1]), ctypes.c_int(B), ctypes.c_int(T), _stream()),
              "kr_energy_norm")
        return buf[1, 0] if squeeze else buf[1]


class FeaturePipeline:
    """Batched, on-device version of the per-item feature code of ``RuslanDataset.__getitem__``
    (data/dataset.py:672-815): peak normalisation, log-mel, pitch and energy aligned to the mel frame count.

    ``FeaturePipeline()(wav, lengths)`` -> dict with ``mel_spec`` (B, 80, T) log-mel, ``pitch`` (B, T), ``energy`` (B, T),
    ``mel_lengths`` (B,); rows are zero beyond each utterance's own ``1 + len // 256`` frames.  Five kernel launches for
    the whole batch; nothing leaves the device."""

    def __init__(self, sample_rate: int = 22050, n_fft: int = 1024, win_length: int = 1024, hop_length: int = 256,
                 n_mels: int = 80, f_min: float = 0.0, f_max: float = 8000.0, pitch_fmin: float = 50.0,
                 pitch_fmax: float = 800.0, device="cuda"):
        self.mel = LogMelSpectrogram(sample_rate, n_fft, win_length, hop_length, n_mels, f_min, f_max, device)
        self.sample_rate, self.hop, self.n_fft = sample_rate, hop_length, n_fft
        self.pitch_fmin, self.pitch_fmax = pitch_fmin, pitch_fmax

    def __call__(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> dict:
        if wav.dim() == 1:
            wav = wav.unsqueeze(0)
        wav = wav.to(self.mel.device, torch.float32).contiguous()
        B, n_max = wav.shape
        if lengths is None:
            lengths = torch.full((B,), n_max, dtype=torch.int64)
        lengths_dev = lengths.to(self.mel.device, torch.int64)
        mel = self.mel(wav, lengths_dev)                                              # (B, 80, T); also pads short clips
        T = mel.shape[2]
        mel_lengths = 1 + torch.clamp(lengths_dev, min=self.n_fft) // self.hop       # dataset.py:687-697
        # the dataset extracts pitch from the peak-normalised audio (:672, :793); the scale cancels in the CMND and in
        # every relative threshold except the 1e-8 / 1e-9 floors, so the normalised waveform is what must be analysed
        peak = torch.empty(B, dtype=torch.float32, device=self.mel.device)
        check(lib().kr_wave_peak(_ptr(wav), _ptr(lengths_dev), _ptr(peak), ctypes.c_int(B), ctypes.c_longlong(n_max),
                                 _stream()), "kr_wave_peak")
        pitch_full = PitchExtractor.extract_pitch(wav / (peak[:, None] + 1e-9), self.sample_rate, self.hop,
                                                  self.pitch_fmin, self.pitch_fmax,
                                                  lengths=torch.clamp(lengths_dev, min=self.n_fft))
        pitch = torch.zeros(B, T, dtype=torch.float32, device=self.mel.device)        # match length to mel frames, :801-806
        n = min(T, pitch_full.shape[1])
        pitch[:, :n] = pitch_full[:, :n]
        pitch = pitch * (torch.arange(T, device=pitch.device)[None, :] < mel_lengths[:, None])
        energy = EnergyExtractor.extract_energy_from_mel(mel, log_domain=False, frames=mel_lengths, channel_major=True,
                                                         exp_input=True)             # :813-815
        return {"mel_spec": mel, "pitch": pitch, "energy": energy, "mel_lengths": mel_lengths}


# ----------------------------------------------------------------------------------------------------------------
# Speed perturbation (data/dataset.py:674-684): torchaudio.functional.resample with torchaudio's defaults, batched
# ----------------------------------------------------------------------------------------------------------------
def resample(waveform: torch.Tensor, orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99,
             lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``torchaudio.functional.resample(waveform, orig_freq, new_freq)`` (Hann-windowed sinc) for (N,) or (B, N) CUDA
    tensors.  ``lengths`` (new, optional): per-row sample counts of a padded batch; rows are zero beyond
    ``ceil(new * len / orig)``.  No filter bank is materialised (csrc/kr_resample_core.cuh)."""
    _need_cuda(waveform, "resample")
    if int(orig_freq) != orig_freq or int(new_freq) != new_freq:
        raise Exception("Frequencies must be of integer type to ensure quality resampling computation.")
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive.")
    squeeze = waveform.dim() == 1
    x = (waveform.unsqueeze(0) if squeeze else waveform).to(torch.float32).contiguous()
    if int(orig_freq) == int(new_freq):
        return waveform
    B, n_max = x.shape
    L = lib()
    L.kr_resample_length.restype = ctypes.c_longlong
    m_max = int(L.kr_resample_length(ctypes.c_longlong(n_max), ctypes.c_int(int(orig_freq)), ctypes.c_int(int(new_freq))))
    if lengths is not None:
        lengths = lengths.to(x.device, torch.int64).contiguous()
    y = torch.empty(B, m_max, dtype=torch.float32, device=x.device)
    check(L.kr_resample(_ptr(x), _ptr(lengths), _ptr(y), ctypes.c_int(B), ctypes.c_longlong(n_max), ctypes.c_longlong(m_max),
                        ctypes.c_int(int(orig_freq)), ctypes.c_int(int(new_freq)), ctypes.c_int(int(lowpass_filter_width)),
                        ctypes.c_float(rolloff), _stream()), "kr_resample")
    return y[0] if squeeze else y


def speed_perturb(waveform: torch.Tensor, factor: float, sample_rate: int = 22050,
                  lengths: Optional[torch.Tensor] = None):
    """The dataset's speed perturbation, data/dataset.py:677-684: resample to ``int(sample_rate * factor)`` and re-normalise
    to unit peak.  Returns (waveform, lengths) — lengths are the resampled sample counts (None in, None out)."""
    if factor == 1.0:
        return waveform, lengths
    new_sr = int(sample_rate * factor)
    y = resample(waveform, sample_rate, new_sr, lengths=lengths)
    y2 = y.unsqueeze(0) if y.dim() == 1 else y
    B, m_max = y2.shape
    new_len = None
    if lengths is not None:
        g = math.gcd(sample_rate, new_sr)
        o, n = sample_rate // g, new_sr // g
        new_len = (lengths.to(y.device, torch.int64) * n + o - 1) // o
    peak = torch.empty(B, dtype=torch.float32, device=y.device)
    check(lib().kr_wave_peak(_ptr(y2), _ptr(new_len), _ptr(peak), ctypes.c_int(B), ctypes.c_longlong(m_max), _stream()),
          "kr_wave_peak")
    out = y2 / (peak[:, None] + 1e-9)
    return (out[0] if y.dim() == 1 else out), new_len


def trailing_trim_end(mel_spec: torch.Tensor, frames: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Frames to keep per utterance after the reference's conservative trailing-silence trim
    (src/kokoro/inference/inference.py:590-621).  ``mel_spec``: (B, T, n_mels) or (T, n_mels) log-mel, already clamped to
    [-11.5, 2] like the reference's input.  Returns an int32 tensor (B,) (or a 0-d tensor) on the device."""
    _need_cuda(mel_spec, "trailing_trim_end")
    squeeze = mel_spec.dim() == 2
    mel = (mel_spec.unsqueeze(0) if squeeze else mel_spec).to(torch.float32).contiguous()
    B, T, M = mel.shape
    if frames is not None:
        frames = frames.to(mel.device, torch.int64).contiguous()
    e = torch.empty(B, T, dtype=torch.float32, device=mel.device)
    t_end = torch.empty(B, dtype=torch.int32, device=mel.device)
    check(lib().kr_energy_frames(_ptr(mel), _ptr(e), ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(M), ctypes.c_int(1),
                                 ctypes.c_int(0), ctypes.c_int(1), _stream()), "kr_energy_frames")
    check(lib().kr_trim_end(_ptr(e), _ptr(frames), _ptr(t_end), ctypes.c_int(B), ctypes.c_int(T), _stream()), "kr_trim_end")
    return t_end[0] if squeeze else t_end
