"""Generates tests/golden/*.npz by importing the LIVE reference from /root/reference/src.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Weights come from oracle.acoustic.seeded_state_dict (construction-order independent), inputs from
oracle.acoustic.synthetic_batch, so fixtures hold only outputs / losses / gradient summaries.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

from oracle import acoustic as oa  # noqa: E402


def _ref_model(cfg: oa.AcousticConfig):
    from kokoro.model.model import KokoroModel
    m = KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim,
                    n_encoder_layers=cfg.n_encoder_layers, n_heads=cfg.n_heads,
                    encoder_ff_dim=cfg.ff_dim, encoder_dropout=0.0, decoder_dropout=0.0,
                    decoder_input_dropout=0.0, n_decoder_layers=cfg.n_decoder_layers,
                    decoder_ff_dim=cfg.ff_dim, max_decoder_seq_len=cfg.max_len,
                    variance_filter_size=cfg.variance_filter, variance_dropout=0.0,
                    n_variance_bins=cfg.n_bins, pitch_min=0.0, pitch_max=1.0, energy_min=0.0,
                    energy_max=1.0, use_stochastic_depth=False, qk_norm=True, ffn_output_norm=True)
    return m


def _ref_losses(cfg, model, outs, batch):
    from kokoro.training.losses import calculate_training_losses
    from kokoro.utils.lengths import average_by_duration
    import logging
    conf = types.SimpleNamespace(duration_loss_weight=cfg.w_dur, stop_token_loss_weight=cfg.w_stop,
                                 pitch_loss_weight=cfg.w_pitch, energy_loss_weight=cfg.w_energy,
                                 verbose=False)
    # criteria exactly as the reference trainer builds them (training/trainer.py:410-436)
    crit = dict(criterion_mel=torch.nn.L1Loss(reduction="none"),
                criterion_duration=torch.nn.HuberLoss(reduction="none", delta=1.0),
                criterion_stop_token=torch.nn.BCEWithLogitsLoss(
                    reduction="none", pos_weight=torch.tensor(cfg.stop_pos_weight)),
                criterion_pitch=torch.nn.HuberLoss(reduction="none", delta=cfg.huber_delta_var),
                criterion_energy=torch.nn.HuberLoss(reduction="none", delta=cfg.huber_delta_var))
    mel, dur, stop, pitch, energy = outs
    return calculate_training_losses(
        device=torch.device("cpu"), config=conf, model=model, average_by_duration=average_by_duration,
        logger=logging.getLogger("golden"), predicted_mel=mel, predicted_log_durations=dur,
        predicted_stop_logits=stop, mel_specs=batch["mel_specs"],
        phoneme_durations=batch["phoneme_durations"], stop_token_targets=batch["stop_token_targets"],
        mel_lengths=batch["mel_lengths"], phoneme_lengths=batch["phoneme_lengths"],
        predicted_pitch=pitch, predicted_energy=energy, pitch_targets=batch["pitches"],
        energy_targets=batch["energies"], **crit)


CASES = {
    # name: (config, batch kwargs)
    "tiny": (oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2,
                               ff_dim=256, variance_filter=64, max_len=1200),
             dict(B=3, P=24, T=150, seed=11, ragged=True)),
    "chunked": (oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1,
                                  ff_dim=128, variance_filter=64, max_len=1200),
                dict(B=2, P=40, T=600, seed=12, ragged=True)),
    "full_width": (oa.AcousticConfig(max_len=1200), dict(B=2, P=32, T=200, seed=13, ragged=True)),
}


def run_case(name):
    cfg, bk = CASES[name]
    batch = oa.synthetic_batch(n_mels=cfg.mel_dim, vocab=cfg.vocab_size, **bk)
    sd = oa.seeded_state_dict(cfg, seed=0)
    model = _ref_model(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    model.train()
    outs = model(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                 batch["stop_token_targets"], pitch_targets=batch["pitches"],
                 energy_targets=batch["energies"], stress_indices=batch["stress_indices"])
    losses = _ref_losses(cfg, model, outs, batch)
    losses[0].backward()
    fix = {f"out_{k}": v.detach().numpy() for k, v in
           zip(("mel", "log_dur", "stop", "pitch", "energy"), outs)}
    fix["losses"] = np.array([float(x) for x in losses], dtype=np.float64)
    names, norms, samples = [], [], []
    for n, p in model.named_parameters():
        names.append(n)
        if p.grad is None:
            norms.append(-1.0)
            samples.append(np.zeros(4, dtype=np.float32))
        else:
            g = p.grad.detach().reshape(-1)
            norms.append(float(g.double().norm()))
            idx = torch.linspace(0, g.numel() - 1, 4).long()
            samples.append(g[idx].numpy())
    fix["grad_names"] = np.array(names)
    fix["grad_norms"] = np.array(norms, dtype=np.float64)
    fix["grad_samples"] = np.stack(samples)
    np.savez_compressed(os.path.join(HERE, f"acoustic_{name}.npz"), **fix)
    # cross-check the oracle right here
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    o_outs = oa.forward_training(sdr, cfg, batch["phoneme_indices"], batch["mel_specs"],
                                 batch["phoneme_durations"], batch["pitches"], batch["energies"],
                                 batch["stress_indices"])
    o_losses = oa.training_losses(cfg, o_outs, batch["mel_specs"], batch["phoneme_durations"],
                                  batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                  batch["mel_lengths"], batch["phoneme_lengths"])
    o_losses[0].backward()
    worst = max(float((a.detach() - b.detach()).abs().max()) for a, b in zip(outs, o_outs))
    gw = 0.0
    for n, p in model.named_parameters():
        og = sdr[n].grad
        if p.grad is None:
            assert og is None or float(og.abs().max()) == 0.0, n
            continue
        gw = max(gw, float((p.grad - og).abs().max() / (p.grad.abs().max() + 1e-12)))
    print(f"{name}: ref vs oracle max|d out|={worst:.3e} max rel d grad={gw:.3e} "
          f"losses ref={[round(float(x), 6) for x in losses]} oracle={[round(float(x), 6) for x in o_losses]}")


if __name__ == "__main__":
    torch.manual_seed(0)
    for name in CASES:
        run_case(name)
