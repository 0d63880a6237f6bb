"""Generates tests/golden/inference.npz by importing the LIVE reference (build container only):
    python tests/golden/make_golden_inference.py
Tiny seeded model in eval() mode; the duration head's bias is raised so that the predicted durations give a
realistic expanded length, the stop head's bias is lowered so that generation runs past the minimum length."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")
from oracle import acoustic as oa  # noqa: E402
from kokoro.model.model import KokoroModel  # noqa: E402

cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256, variance_filter=64,
                        max_len=1200)
sd = oa.seeded_state_dict(cfg, seed=4)
sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([1.2])
sd["stop_token_predictor.bias"] = torch.tensor([-1.0])
m = KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim, n_encoder_layers=2, n_heads=2,
                encoder_ff_dim=cfg.ff_dim, encoder_dropout=0.1, decoder_dropout=0.1, decoder_input_dropout=0.1,
                n_decoder_layers=2, decoder_ff_dim=cfg.ff_dim, max_decoder_seq_len=cfg.max_len,
                variance_filter_size=cfg.variance_filter, variance_dropout=0.1, n_variance_bins=cfg.n_bins, pitch_min=0.0,
                pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=True, qk_norm=True, ffn_output_norm=True)
m.load_state_dict(sd, strict=True)
m.eval()
g = torch.Generator().manual_seed(8)
idx = torch.randint(1, cfg.vocab_size, (1, 21), generator=g)
stress = torch.randint(0, 3, (1, 21), generator=g)
mel = m.forward_inference(idx, stress_indices=stress)
idx2 = torch.randint(1, cfg.vocab_size, (2, 13), generator=g)
idx2[1, 9:] = 0                                   # padded tail in a batch of two
mel2 = m.forward_inference(idx2, stress_indices=None, stop_threshold=0.45)
out = os.path.join(HERE, "inference.npz")
np.savez_compressed(out, idx=idx.numpy(), stress=stress.numpy(), mel=mel.numpy(), idx2=idx2.numpy(), mel2=mel2.numpy(),
                    dur_bias=np.array(1.2), stop_bias=np.array(-1.0), seed=np.array(4))
print("wrote", out, mel.shape, mel2.shape)
