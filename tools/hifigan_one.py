"""One eager HiFi-GAN forward at config-5 size (profiled under ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
gen = HiFiGANGenerator(HiFiGANConfig.get_default_config(), device="cuda", use_graphs=False)
mel = (torch.randn(16, 80, 800) * 2 - 5).cuda()
for _ in range(2):
    a = gen(mel)
torch.cuda.synchronize()
print(a.shape, float(a.abs().max()))
