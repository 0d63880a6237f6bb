"""The fp32 oracle against the INSTALLED reference (baseline/_ref), side by side on edge batches the committed fixtures do
not hold: a batch of one, the shortest sequences the model accepts, mel lengths on / around the VariancePredictor's 512-frame
chunk boundary (model/variance_predictor.py:77-87), one utterance much shorter than its batch.  Outputs, losses and every
gradient norm (rtol 1e-4: both sides are fp32 on the CPU).  Skipped where baseline/_ref is absent (the GPU box never needs
it: the oracle is what travels)."""
import logging
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def _reference():
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)


def _ref_model(cfg):
    from kokoro.model.model import KokoroModel
    return KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim,
                       n_encoder_layers=cfg.n_encoder_layers, n_heads=cfg.n_heads, encoder_ff_dim=cfg.ff_dim,
                       encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0,
                       n_decoder_layers=cfg.n_decoder_layers, decoder_ff_dim=cfg.ff_dim, max_decoder_seq_len=cfg.max_len,
                       variance_filter_size=cfg.variance_filter, variance_dropout=0.0, n_variance_bins=cfg.n_bins,
                       pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=False,
                       qk_norm=True, ffn_output_norm=True)


def _ref_losses(cfg, model, outs, batch):
    from kokoro.training.losses import calculate_training_losses
    from kokoro.utils.lengths import average_by_duration
    conf = types.SimpleNamespace(duration_loss_weight=cfg.w_dur, stop_token_loss_weight=cfg.w_stop,
                                 pitch_loss_weight=cfg.w_pitch, energy_loss_weight=cfg.w_energy, verbose=False)
    crit = dict(criterion_mel=torch.nn.L1Loss(reduction="none"),
                criterion_duration=torch.nn.HuberLoss(reduction="none", delta=1.0),
                criterion_stop_token=torch.nn.BCEWithLogitsLoss(reduction="none", pos_weight=torch.tensor(cfg.stop_pos_weight)),
                criterion_pitch=torch.nn.HuberLoss(reduction="none", delta=cfg.huber_delta_var),
                criterion_energy=torch.nn.HuberLoss(reduction="none", delta=cfg.huber_delta_var))
    mel, dur, stop, pitch, energy = outs
    return calculate_training_losses(
        device=torch.device("cpu"), config=conf, model=model, average_by_duration=average_by_duration,
        logger=logging.getLogger("edges"), predicted_mel=mel, predicted_log_durations=dur, predicted_stop_logits=stop,
        mel_specs=batch["mel_specs"], phoneme_durations=batch["phoneme_durations"],
        stop_token_targets=batch["stop_token_targets"], mel_lengths=batch["mel_lengths"],
        phoneme_lengths=batch["phoneme_lengths"], predicted_pitch=pitch, predicted_energy=energy,
        pitch_targets=batch["pitches"], energy_targets=batch["energies"], **crit)


EDGES = {
    "batch of one": dict(B=1, P=9, T=40, seed=21, ragged=False),
    "shortest sequences": dict(B=2, P=2, T=4, seed=22, ragged=False),
    "exactly one predictor chunk": dict(B=2, P=64, T=512, seed=23, ragged=False),
    "one frame past the chunk": dict(B=2, P=64, T=513, seed=24, ragged=True),
    "two chunks and a tail": dict(B=1, P=130, T=1100, seed=25, ragged=False),
    "ragged with a very short utterance": dict(B=4, P=30, T=220, seed=26, ragged=True),
}


@pytest.mark.parametrize("label", list(EDGES))
def test_oracle_equals_the_installed_reference_on_edge_batches(label):
    _reference()
    from oracle import acoustic as oa
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=2, ff_dim=128, variance_filter=64,
                            max_len=1200)
    batch = oa.synthetic_batch(n_mels=cfg.mel_dim, vocab=cfg.vocab_size, **EDGES[label])
    if label == "ragged with a very short utterance":            # utterance 1: 3 phonemes, their frames only
        batch["phoneme_lengths"][1] = 3
        batch["phoneme_durations"][1, 3:] = 0
        n = int(batch["phoneme_durations"][1].sum())
        batch["mel_lengths"][1] = n
        for k in ("mel_specs", "pitches", "energies", "stop_token_targets"):
            batch[k][1, n:] = 0
        batch["stop_token_targets"][1] = 0
        batch["stop_token_targets"][1, :n] = oa.build_stop_token_targets(n)
    sd = oa.seeded_state_dict(cfg, seed=0)
    model = _ref_model(cfg)
    model.load_state_dict(sd, strict=True)
    model.train()
    want = model(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                 pitch_targets=batch["pitches"], energy_targets=batch["energies"], stress_indices=batch["stress_indices"])
    want_losses = _ref_losses(cfg, model, want, batch)
    want_losses[0].backward()
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    got = oa.forward_training(sdr, cfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                              batch["pitches"], batch["energies"], batch["stress_indices"])
    got_losses = oa.training_losses(cfg, got, batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                                    batch["pitches"], batch["energies"], batch["mel_lengths"], batch["phoneme_lengths"])
    got_losses[0].backward()
    for key, g, w in zip(("mel", "log_dur", "stop", "pitch", "energy"), got, want):
        assert g.shape == w.shape, (label, key, g.shape, w.shape)
        assert torch.allclose(g.detach(), w.detach(), rtol=1e-4, atol=2e-5), (label, key, float((g - w).abs().max()))
    for i, (g, w) in enumerate(zip(got_losses, want_losses)):
        assert abs(float(g) - float(w)) <= 1e-5 * abs(float(w)) + 1e-6, (label, i, float(g), float(w))
    for name, p in model.named_parameters():
        g = sdr[name].grad
        if p.grad is None or float(p.grad.abs().max()) == 0.0:
            assert g is None or float(g.abs().max()) <= 1e-7, (label, name)
            continue
        wn = float(p.grad.double().norm())
        assert abs(float(g.double().norm()) - wn) <= 1e-4 * wn + 1e-7, (label, name, float(g.double().norm()), wn)
