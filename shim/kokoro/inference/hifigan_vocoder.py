"""HiFi-GAN generator (reference src/kokoro/inference/hifigan_vocoder.py:31-271)."""
from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator, load_hifigan_model  # noqa: F401
