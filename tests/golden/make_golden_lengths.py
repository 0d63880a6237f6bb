"""Generates tests/golden/lengths.npz by importing the LIVE reference (run in the build container only):
    PYTHONPATH=/root/reference/src python tests/golden/make_golden_lengths.py
Inputs are seeded; outputs are the reference's own `vectorized_expand_tokens` and `length_regulate`."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from kokoro.utils.lengths import length_regulate, vectorized_expand_tokens  # noqa: E402

g = torch.Generator().manual_seed(7)
B, P, D = 4, 37, 16
enc = torch.randn(B, P, D, generator=g)
dur = torch.randint(0, 9, (B, P), generator=g)
dur[1, 20:] = 0
dur[2] = 0                      # all-zero row
dur[3, ::4] = -3                # negative durations
pad = torch.zeros(B, P, dtype=torch.bool)
pad[0, 30:] = True
pad[1, 20:] = True
pad[3] = True                   # fully padded sample
exp_a = vectorized_expand_tokens(enc, dur)
exp_b = vectorized_expand_tokens(enc, dur, max_len=120)
exp_c = vectorized_expand_tokens(enc[..., 0], dur)
fb_out, fb_mask = length_regulate(enc, dur.float(), pad)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lengths.npz")
np.savez_compressed(out, enc=enc.numpy(), dur=dur.numpy(), pad=pad.numpy(), exp_a=exp_a.numpy(), exp_b=exp_b.numpy(),
                    exp_c=exp_c.numpy(), fb_out=fb_out.numpy(), fb_mask=fb_mask.numpy())
print("wrote", out, exp_a.shape, exp_b.shape, fb_out.shape)
