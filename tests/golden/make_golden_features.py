"""Generates tests/golden/features.npz by importing the LIVE reference (build container only):
    python tests/golden/make_golden_features.py
Synthetic utterances: harmonic tones with vibrato and glides, noise bursts, silences and short gaps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from kokoro.model.variance_predictor import EnergyExtractor, PitchExtractor  # noqa: E402

SR = 22050
rng = np.random.default_rng(5)


def utterance(seconds, f0_start, f0_end, vib=0.0):
    n = int(SR * seconds)
    t = np.arange(n) / SR
    f0 = np.linspace(f0_start, f0_end, n) * (1.0 + vib * np.sin(2 * np.pi * 5.5 * t))
    ph = 2 * np.pi * np.cumsum(f0) / SR
    x = sum(np.sin(k * ph) / k for k in range(1, 6))
    env = np.ones(n)
    for s in rng.integers(0, n - 4000, 4):            # silences / gaps of different lengths
        env[s:s + int(rng.integers(300, 4000))] = 0.0
    x = x * env * 0.3 + rng.normal(0, 0.003, n)
    noise = int(rng.integers(0, n - 6000))
    x[noise:noise + 5000] = rng.normal(0, 0.2, 5000)   # unvoiced burst
    return x.astype(np.float32)


waves = [utterance(1.6, 110, 180, 0.02), utterance(1.6, 320, 220, 0.0), utterance(1.6, 90, 95, 0.05),
         utterance(1.6, 600, 750, 0.01)]
wav = torch.from_numpy(np.stack(waves))
pitch = PitchExtractor.extract_pitch(wav, sample_rate=SR, hop_length=256, fmin=50.0, fmax=800.0)
short = PitchExtractor.extract_pitch(wav[0, :1500], sample_rate=SR, hop_length=256)
g = torch.Generator().manual_seed(3)
mel = torch.randn(3, 140, 80, generator=g) * 2.0 - 5.0
mel[1, 100:] -= 4.0
e_log = EnergyExtractor.extract_energy_from_mel(mel)
e_lin = EnergyExtractor.extract_energy_from_mel(mel.exp(), log_domain=False)
e_short = EnergyExtractor.extract_energy_from_mel(mel[:, :2])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "features.npz")
np.savez_compressed(out, wav=wav.numpy(), pitch=pitch.numpy(), pitch_short=short.numpy(), mel=mel.numpy(),
                    e_log=e_log.numpy(), e_lin=e_lin.numpy(), e_short=e_short.numpy())
print("wrote", out, pitch.shape, short.shape, "voiced fraction", float((pitch > 0).float().mean()))
