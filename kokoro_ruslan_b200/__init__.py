"""kokoro_ruslan_b200 — Blackwell-native (sm_100a) hot path of igorshmukler/kokoro-ruslan.

Acoustic-model training step + HiFi-GAN vocoder inference behind the reference's own Python
surfaces; compute is hand-written CUDA in libkokoro_b200.so reached through a C ABI
(include/kokoro_b200.h).
"""
__version__ = "0.1.0"
