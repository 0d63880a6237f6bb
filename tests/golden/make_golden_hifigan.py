"""Generates tests/golden/hifigan_small.npz from the LIVE reference HiFiGANGenerator
(/root/reference/src; build container only).  Weights: oracle.hifigan.seeded_state_dict, loaded into
the reference module with strict=True (so the key set / shapes are also checked)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import hifigan as oh  # noqa: E402


def main():
    from kokoro.inference.hifigan_vocoder import HiFiGANConfig, HiFiGANGenerator
    cfg = oh.HifiConfig()
    ref = HiFiGANGenerator(HiFiGANConfig.get_default_config()).eval()
    sd = oh.seeded_state_dict(cfg, seed=0)
    assert [k for k in ref.state_dict().keys()] == list(sd.keys()), "state-dict key order differs"
    ref.load_state_dict(sd, strict=True)
    out = {}
    for name, (B, T, seed) in {"a": (2, 37, 1), "b": (1, 64, 2)}.items():
        mel = oh.synthetic_mel(B, T, seed)
        with torch.no_grad():
            y = ref(mel)
            y_t = ref(mel.transpose(1, 2).contiguous())      # (B,T,80) layout
        assert torch.equal(y, y_t)
        mine = oh.generator_forward(sd, cfg, mel)
        print(name, tuple(y.shape), "oracle vs reference max abs diff", float((mine - y).abs().max()))
        out[f"audio_{name}"] = y.numpy()
        out[f"shape_{name}"] = np.array([B, T, seed])
    np.savez_compressed(os.path.join(HERE, "hifigan_small.npz"), **out)


if __name__ == "__main__":
    main()
