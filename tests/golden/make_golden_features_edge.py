"""Generates tests/golden/features_edge.npz by importing the LIVE reference (build container only):
    python tests/golden/make_golden_features_edge.py
Edge inputs of PitchExtractor.extract_pitch: silence, DC, a single sample, white noise, an impulse, tones at both ends of
the 50-800 Hz range; EnergyExtractor on 1-5 frames."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from kokoro.model.variance_predictor import EnergyExtractor, PitchExtractor  # noqa: E402

rng = np.random.default_rng(0)
t = np.arange(15000) / 22050
cases = {"zeros": np.zeros(8000, np.float32), "dc": np.full(8000, 0.3, np.float32), "n1": np.array([0.5], np.float32),
         "noise": rng.normal(0, 0.1, 12000).astype(np.float32), "impulse": np.eye(1, 9000, 4000, dtype=np.float32)[0],
         "tone60": (0.5 * np.sin(2 * np.pi * 60 * t)).astype(np.float32),
         "tone790": (0.5 * np.sin(2 * np.pi * 790 * t)).astype(np.float32)}
out = {}
for k, x in cases.items():
    out[f"wav_{k}"] = x
    out[f"pitch_{k}"] = PitchExtractor.extract_pitch(torch.from_numpy(x)).numpy()
for T in (1, 2, 3, 5):
    m = torch.full((T, 80), -5.0)
    m[0] += 1
    out[f"mel_{T}"] = m.numpy()
    out[f"energy_{T}"] = EnergyExtractor.extract_energy_from_mel(m, log_domain=True).numpy()
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "features_edge.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
