"""Generates tests/golden/resample.npz with the LIVE torchaudio (build container only):
    python tests/golden/make_golden_resample.py
For each speed factor of the reference's perturbation range (data/dataset.py:674-684: new_sr = int(22050 * factor)):
torchaudio.functional.resample's output, and — at sampled positions — the float64-accumulated application of
torchaudio's own filter bank (_get_sinc_resample_kernel), which is the exact value of the algorithm; torchaudio's strided
float32 conv1d over its 11 k-tap rows deviates from that by up to 2e-4 for ratios that do not reduce."""
import math
import os

import numpy as np
import torch
import torchaudio
import torchaudio.functional.functional as F

SR = 22050
g = torch.Generator().manual_seed(0)
t = torch.arange(12000) / SR
x = torch.stack([0.3 * torch.sin(2 * math.pi * 220 * t) + 0.05 * torch.randn(12000, generator=g),
                 0.2 * torch.randn(12000, generator=g)])
out = {"x": x.numpy(), "factors": np.array([0.9, 0.93, 1.07, 1.1])}
rng = np.random.default_rng(0)
for f in out["factors"]:
    new = int(SR * float(f))
    y = torchaudio.functional.resample(x, SR, new)
    gg = math.gcd(SR, new)
    orig, neu = SR // gg, new // gg
    kernel, width = F._get_sinc_resample_kernel(SR, new, gg)
    xp = torch.nn.functional.pad(x, (width, width + orig)).double()
    js = np.sort(rng.integers(0, y.shape[1], 400))
    exact = np.zeros((2, len(js)))
    for b in range(2):
        for n, j in enumerate(js):
            i, p = divmod(int(j), neu)
            seg = xp[b, i * orig:i * orig + kernel.shape[-1]]
            exact[b, n] = float((seg * kernel[p, 0, :len(seg)].double()).sum())
    key = f"{new}"
    out[f"y_{key}"], out[f"js_{key}"], out[f"exact_{key}"] = y.numpy(), js, exact
    print(f, new, y.shape, float(np.abs(y.numpy()[:, js] - exact).max()))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resample.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path))
