"""Generates tests/golden/average_by_duration.npz by importing the LIVE reference (build container only):
    python tests/golden/make_golden_average.py
Seeded cases for kokoro.utils.lengths.average_by_duration incl. its quirks: durations summing past the last frame,
durations that leave trailing frames uncovered (they are averaged into token 0), negative / zero durations, masks."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from kokoro.utils.lengths import average_by_duration  # noqa: E402

g = torch.Generator().manual_seed(3)
out = {}
for k, (B, P, T, hi) in enumerate(((4, 37, 120, 9), (3, 20, 50, 9), (2, 128, 800, 12), (3, 9, 30, 2))):
    v = torch.randn(B, T, generator=g)
    d = torch.randint(-2, hi, (B, P), generator=g)
    m = torch.rand(B, P, generator=g) < 0.2
    out[f"v{k}"], out[f"d{k}"], out[f"m{k}"] = v.numpy(), d.numpy(), m.numpy()
    out[f"a{k}"] = average_by_duration(v, d).numpy()
    out[f"am{k}"] = average_by_duration(v, d, m).numpy()
    print(k, v.shape, d.shape, int(d.clamp(min=0).sum(1).max()))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "average_by_duration.npz")
np.savez_compressed(path, **out)
print("wrote", path)
