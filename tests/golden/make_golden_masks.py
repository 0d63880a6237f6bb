"""Generates tests/golden/acoustic_masks.npz: the LIVE reference forward (train mode, dropout 0) with EXPLICIT
text_padding_mask / mel_padding_mask arguments (model/model.py:586-589, 648-654) — the pin of the oracle's and the product's
explicit-mask path.  Run in the build container only:  python tests/golden/make_golden_masks.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import CASES, _ref_model  # noqa: E402  (also puts /root/reference/src on sys.path)
from oracle import acoustic as oa  # noqa: E402


def masks_for(batch):
    """Text mask: the default (id 0) PLUS the last real token of every utterance; mel mask: frames >= mel_length."""
    idx = batch["phoneme_indices"]
    B, P = idx.shape
    text = idx == 0
    for b in range(B):
        text[b, int(batch["phoneme_lengths"][b]) - 1] = True
    T = batch["mel_specs"].shape[1]
    mel = torch.arange(T).unsqueeze(0) >= batch["mel_lengths"].unsqueeze(1)
    return text, mel


if __name__ == "__main__":
    cfg, bk = CASES["tiny"]
    batch = oa.synthetic_batch(n_mels=cfg.mel_dim, vocab=cfg.vocab_size, **bk)
    sd = oa.seeded_state_dict(cfg, seed=0)
    model = _ref_model(cfg)
    model.load_state_dict(sd, strict=True)
    model.train()
    text, mel = masks_for(batch)
    with torch.no_grad():
        outs = model(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                     pitch_targets=batch["pitches"], energy_targets=batch["energies"], text_padding_mask=text,
                     mel_padding_mask=mel, stress_indices=batch["stress_indices"])
        base = model(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                     pitch_targets=batch["pitches"], energy_targets=batch["energies"], stress_indices=batch["stress_indices"])
        o = oa.forward_training(sd, cfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                batch["pitches"], batch["energies"], batch["stress_indices"], text_padding_mask=text,
                                mel_padding_mask=mel)
    names = ("mel", "log_dur", "stop", "pitch", "energy")
    fix = {f"out_{k}": v.numpy() for k, v in zip(names, outs)}
    np.savez_compressed(os.path.join(HERE, "acoustic_masks.npz"), **fix)
    print("masks change the outputs by", max(float((a - b).abs().max()) for a, b in zip(outs, base)))
    print("oracle vs live reference:", max(float((a - b).abs().max()) for a, b in zip(outs, o)))
