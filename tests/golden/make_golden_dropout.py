"""Generates tests/golden/acoustic_dropout.npz by importing the LIVE reference (build container only):
    python tests/golden/make_golden_dropout.py
The reference model runs in train() mode with ALL its dropouts and stochastic depth on, after torch.manual_seed(SEED).
The oracle (oracle.acoustic.forward_training with the TorchDropout callback) must reproduce these outputs when it
draws from the same torch RNG stream: same sites, same order, same shapes, same semantics."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/src")
from oracle import acoustic as oa  # noqa: E402
from kokoro.model.model import KokoroModel  # noqa: E402

SEED = 123
P = dict(enc=0.15, dec=0.2, din=0.15, var=0.1, sd=0.5)
cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256, variance_filter=64,
                        max_len=1200)
batch = oa.synthetic_batch(n_mels=cfg.mel_dim, vocab=cfg.vocab_size, B=3, P=24, T=150, seed=11, ragged=True)
sd = oa.seeded_state_dict(cfg, seed=0)
m = KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim, n_encoder_layers=2, n_heads=2,
                encoder_ff_dim=cfg.ff_dim, encoder_dropout=P["enc"], decoder_dropout=P["dec"],
                decoder_input_dropout=P["din"], n_decoder_layers=2, decoder_ff_dim=cfg.ff_dim,
                max_decoder_seq_len=cfg.max_len, variance_filter_size=cfg.variance_filter, variance_dropout=P["var"],
                n_variance_bins=cfg.n_bins, pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0,
                use_stochastic_depth=True, stochastic_depth_rate=P["sd"], qk_norm=True, ffn_output_norm=True)
m.load_state_dict(sd, strict=True)
m.train()
torch.manual_seed(SEED)
outs = m(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
         pitch_targets=batch["pitches"], energy_targets=batch["energies"], stress_indices=batch["stress_indices"])
fix = {f"out_{k}": v.detach().numpy() for k, v in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs)}
fix["seed"] = np.array(SEED)
fix["probs"] = np.array([P["enc"], P["dec"], P["din"], P["var"], P["sd"]])
fix["torch_version"] = np.array(torch.__version__)
np.savez_compressed(os.path.join(HERE, "acoustic_dropout.npz"), **fix)
print("wrote acoustic_dropout.npz", {k: v.shape for k, v in fix.items() if hasattr(v, "shape")})
