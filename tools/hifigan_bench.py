"""HiFi-GAN vocoder inference throughput (BASELINE config 5: 16 mels x 800 frames -> 22.05 kHz audio)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(B=16, T=800, steps=20, warmup=3, e2e=True):
    from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
    gen = HiFiGANGenerator(HiFiGANConfig.get_default_config(), device="cuda")
    g = torch.Generator().manual_seed(0)
    mel_host = (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).pin_memory()
    mel_dev = mel_host.cuda()
    for _ in range(max(3, warmup)):
        gen(mel_dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gen(mel_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out_host = torch.empty(B, 1, T * 256).pin_memory()
    e0.record()
    for _ in range(steps):
        a = gen(mel_host)
        out_host.copy_(a, non_blocking=True)
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / steps
    samples = B * T * 256
    flops = 614.1e6 * B * T
    return {"metric": "hifigan_audio_samples_per_sec", "value": samples / (ms * 1e-3), "unit": "samples/s",
            "ms_per_batch": ms, "batch": B, "frames": T, "tflops_algorithmic": flops / (ms * 1e-3) / 1e12,
            "e2e": {"value": samples / (ms_e2e * 1e-3), "ms_per_batch": ms_e2e,
                    "h2d_bytes_per_step": mel_host.numel() * 4, "d2h_bytes_per_step": samples * 4},
            "launches_per_forward": gen.launches_last_forward}


if __name__ == "__main__":
    print(json.dumps(run()))
