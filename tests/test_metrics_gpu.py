"""kr_val_metrics on the device against the literal reference loop, and through
TrainStep.eval_losses(metrics=...).  (The kernel body is also verified by host emulation, tests/test_metrics_emu_cpu.py.)"""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_kernel_matches_reference_loop():
    from kokoro_ruslan_b200 import ops
    from test_metrics_emu_cpu import reference_loop
    g = torch.Generator().manual_seed(0)
    batches = []
    for B, T, lens in ((4, 50, [50, 31, 0, 7]), (1, 20, [20]), (8, 800, [800, 640, 800, 33, 512, 799, 1, 400])):
        mel = torch.randn(B, T, 80, generator=g) * 2 - 5
        mel_pred = mel + 0.3 * torch.randn(B, T, 80, generator=g)
        pitch, pitch_pred = torch.rand(B, T, generator=g), torch.rand(B, T, generator=g)
        batches.append((mel_pred, mel, pitch_pred, pitch, torch.tensor(lens, dtype=torch.int64)))
    acc = ops.zero_(torch.empty(ops.val_metrics_acc_floats(), device="cuda"))
    for b in batches:
        mp, m, pp, p, lens = (t.cuda() for t in b)
        ops.val_metrics(mp, m, pp, p, lens, acc)
    want_sc, want_f0 = reference_loop(batches)
    a = acc.cpu()
    assert a[1] == 3 and a[3] == 3
    assert float(a[0] / a[1]) == pytest.approx(want_sc, rel=1e-4)
    assert float(a[2] / a[3]) == pytest.approx(want_f0, rel=1e-4)


def test_eval_losses_accumulates_metrics():
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    from oracle import acoustic as oa
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    cfg = ModelConfig(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                      n_encoder_layers=2, n_heads=2, encoder_ff_dim=256, n_decoder_layers=2, decoder_ff_dim=256,
                      max_decoder_seq_len=1200, variance_filter_size=64, n_variance_bins=ocfg.n_bins)
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=10), device="cuda:0", use_graphs=False)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    ts.load_state_dict(sd)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    acc = ts.new_val_metrics()
    plain = ts.eval_losses(batch, use_ema=False)
    with_m = ts.eval_losses(batch, use_ema=False, metrics=acc)
    assert torch.equal(plain, with_m)
    got = ts.read_val_metrics(acc)
    outs = oa.forward_training(sd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                               batch["pitches"], batch["energies"], batch["stress_indices"])
    from test_metrics_emu_cpu import reference_loop
    want_sc, want_f0 = reference_loop([(outs[0].detach(), batch["mel_specs"], outs[3].detach(), batch["pitches"],
                                        batch["mel_lengths"])])
    assert got["val_spectral_convergence"] == pytest.approx(want_sc, rel=2e-2)
    assert got["val_f0_rmse"] == pytest.approx(want_f0, rel=2e-2)
