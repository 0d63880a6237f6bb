"""tests/golden/melstft.npz: torchaudio's OWN MelSpectrogram (configured exactly as the reference
data/dataset.py:162-178) + log on seeded waveforms.  Build container only."""
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import melstft as om  # noqa: E402


def waveform(n, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n) / 22050.0
    x = 0.4 * torch.sin(2 * torch.pi * 220.0 * t) + 0.2 * torch.sin(2 * torch.pi * 1930.0 * t * (1 + 0.1 * t))
    return (x + 0.05 * torch.randn(n, generator=g)) * 0.7


def main():
    tr = torchaudio.transforms.MelSpectrogram(sample_rate=22050, n_fft=1024, n_mels=80, hop_length=256, win_length=1024,
                                              f_min=0.0, f_max=8000.0, power=2.0, normalized=False,
                                              window_fn=torch.hann_window)
    out = {}
    for name, (n, seed) in {"a": (22050, 1), "b": (5000, 2), "c": (256 * 37, 3)}.items():
        x = waveform(n, seed)
        xn = x / (x.abs().max() + 1e-9)
        ref = torch.log(tr(xn.unsqueeze(0)).squeeze(0) + 1e-9)
        mine = om.log_mel(x.numpy())
        print(name, tuple(ref.shape), "oracle vs torchaudio max abs diff", float(np.abs(mine - ref.numpy()).max()))
        out[f"wav_{name}"] = x.numpy()
        out[f"mel_{name}"] = ref.numpy()
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm=None, mel_scale="htk")
    print("filterbank max abs diff", float(np.abs(om.mel_filterbank() - fb.numpy()).max()))
    np.savez_compressed(os.path.join(HERE, "melstft.npz"), **out)


if __name__ == "__main__":
    main()
