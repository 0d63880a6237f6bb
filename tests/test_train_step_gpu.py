"""Full optimizer steps through TrainStep (CUDA graphs on) vs the CPU oracle step.

Losses: bf16 path within 1e-2 relative of the fp32 oracle at every step.  Weights: AdamW
normalises the gradient, so bf16 gradient noise can flip the update sign of near-zero-gradient
elements; the gate is therefore on the relative L2 error of the accumulated update."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tiny():
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.params import ModelConfig
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    cfg = ModelConfig(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                      n_encoder_layers=ocfg.n_encoder_layers, n_heads=ocfg.n_heads, encoder_ff_dim=ocfg.ff_dim,
                      n_decoder_layers=ocfg.n_decoder_layers, decoder_ff_dim=ocfg.ff_dim,
                      max_decoder_seq_len=ocfg.max_len, variance_filter_size=ocfg.variance_filter,
                      n_variance_bins=ocfg.n_bins)
    return ocfg, cfg


@pytest.mark.parametrize("graphs", [False, True])
def test_train_steps_match_cpu_oracle(graphs):
    from oracle import acoustic as oa
    from oracle.train_step import CpuTrainStep
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    ocfg, cfg = _tiny()
    sd = oa.seeded_state_dict(ocfg, seed=0)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    lr = 1e-3
    ts = TrainStep(cfg, OptimConfig(learning_rate=lr, ema_decay=0.9), ScheduleConfig(total_steps=1000, use_warmup=False,
                                                                                       pct_start=0.5, max_lr_multiplier=1.0),
                   device="cuda", use_graphs=graphs)
    ts.load_state_dict(sd)
    ref = CpuTrainStep(ocfg, sd, lr=lr, ema_decay=0.9)
    pinned = {k: v.pin_memory() for k, v in batch.items()}
    n_steps = 4
    for k in range(n_steps):
        base_lr = ts.sched.lrs()[2]          # group 2 has multiplier 1
        ref.set_lr(base_lr)
        want = ref.train_step(batch)
        got = ts.train_step(pinned).cpu().tolist()
        for g, w in zip(got, want):
            assert abs(g - w) <= 1e-2 * abs(w) + 1e-4, (k, got, want)
    torch.cuda.synchronize()
    ctrl = ts.opt.read_ctrl()
    assert ctrl["step"] == n_steps and ctrl["skip"] == 0
    num = den = 0.0
    mine = ts.store.state_dict()
    for n in ts.store.order:
        d_ref = ref.sd[n].detach() - sd[n]
        d_got = mine[n].float().cpu() - sd[n]
        num += float((d_got - d_ref).pow(2).sum())
        den += float(d_ref.pow(2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5
    ema = ts.store.state_dict(ts.store.ema)
    worst = max(float((ema[n].float().cpu() - ref.ema[n]).abs().max()) for n in ts.store.order)
    assert worst < 5 * lr * n_steps, worst


def test_adaptive_long_sequence_scalars_reach_device():
    """T > 1400 frames => loss scale 1/r and clip 0.5/sqrt(r) (reference trainer.py:2218-2242)."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep, adaptive_stabilisation
    ocfg, cfg = _tiny()
    cfg.max_decoder_seq_len = 2100
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=10), device="cuda", use_graphs=False)
    ts.store.init_default(seed=0)
    batch = oa.synthetic_batch(B=1, P=64, T=1600, seed=5)
    ts.train_step({k: v.pin_memory() for k, v in batch.items()})
    torch.cuda.synchronize()
    scale, clip = adaptive_stabilisation(1600, int(batch["phoneme_durations"].max()), 1.5)
    assert abs(scale - 1400 / 1600) < 1e-6 and abs(clip - 0.5 / (1600 / 1400) ** 0.5) < 1e-6
    assert abs(float(ts.loss_scale.cpu()) - scale) < 1e-6
    assert abs(ts.opt.read_ctrl()["clip_used"] - clip) < 1e-6


@pytest.mark.parametrize("graphs", [False, True])
def test_gradient_accumulation_window_matches_cpu_oracle(graphs):
    """gradient_accumulation_steps = 2 (the reference default, training/config.py): zero_grad at the window
    start, each micro-batch's loss scaled by 1/divisor, ONE optimizer step at the window end
    (trainer.py:2258-2294, 2341-2343).  Two different batch shapes per window (two graph variants each)."""
    from oracle import acoustic as oa
    from oracle.train_step import CpuTrainStep
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    ocfg, cfg = _tiny()
    sd = oa.seeded_state_dict(ocfg, seed=0)
    b0 = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    b1 = oa.synthetic_batch(B=2, P=32, T=130, seed=12, ragged=True)
    lr = 1e-3
    ts = TrainStep(cfg, OptimConfig(learning_rate=lr, ema_decay=0.9),
                   ScheduleConfig(total_steps=1000, use_warmup=False, pct_start=0.5), device="cuda", use_graphs=graphs)
    ts.load_state_dict(sd)
    ref = CpuTrainStep(ocfg, sd, lr=lr, ema_decay=0.9)
    n_windows = 3
    for k in range(n_windows):
        ref.set_lr(ts.sched.lrs()[2])
        want = []
        for i, b in enumerate((b0, b1)):
            # CpuTrainStep.fwd_bwd clears .grad: accumulate by hand like autograd would
            saved = {n: (ref.sd[n].grad.clone() if ref.sd[n].grad is not None else None) for n in ref.names} if i else None
            _, losses = ref.fwd_bwd(b, loss_scale=0.5)
            if saved is not None:
                for n in ref.names:
                    if saved[n] is not None:
                        ref.sd[n].grad = saved[n] if ref.sd[n].grad is None else ref.sd[n].grad + saved[n]
            want.append([float(x.detach()) for x in losses])
        ref.optimizer_step()
        got = [l.cpu().tolist() for l in ts.train_window([b0, b1])]
        for gw, ww in zip(got, want):
            for g, w in zip(gw, ww):
                assert abs(g - w) <= 1e-2 * abs(w) + 1e-4, (k, got, want)
    torch.cuda.synchronize()
    ctrl = ts.opt.read_ctrl()
    assert ctrl["step"] == n_windows and ctrl["skip"] == 0
    assert ts.sched.current_optimizer_step == n_windows
    num = den = 0.0
    mine = ts.store.state_dict()
    for n in ts.store.order:
        d_ref = ref.sd[n].detach() - sd[n]
        d_got = mine[n].float().cpu() - sd[n]
        num += float((d_got - d_ref).pow(2).sum())
        den += float(d_ref.pow(2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5


def test_validation_forward_uses_ema_weights_and_no_dropout():
    """eval_losses = the reference's validate_epoch forward: EMA weights, eval mode (trainer.py:1771-1824)."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.engine import DropoutConfig
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    ocfg, cfg = _tiny()
    sd = oa.seeded_state_dict(ocfg, seed=0)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    ts = TrainStep(cfg, OptimConfig(learning_rate=1e-3, ema_decay=0.5), ScheduleConfig(total_steps=100, use_warmup=False),
                   device="cuda", use_graphs=True, dropout=DropoutConfig.reference_training())
    ts.load_state_dict(sd)
    for _ in range(3):
        ts.train_step(batch)
    rng_step = int(ts.engine.drop_state[1])
    params_before = ts.store.params.clone()
    got = [ts.eval_losses(batch, use_ema=e).cpu() for e in (True, False, True)]
    assert torch.equal(got[0], got[2])                                  # deterministic: no dropout in eval
    assert not torch.allclose(got[0], got[1], rtol=1e-4)                # EMA weights differ from the live weights
    assert int(ts.engine.drop_state[1]) == rng_step and torch.equal(ts.store.params, params_before)
    for use_ema, g in zip((True, False), got[:2]):
        buf = ts.store.ema if use_ema else ts.store.params
        wsd = {k: v.float().cpu() for k, v in ts.store.state_dict(buf).items()}
        for k in oa.BUFFER_KEYS:
            wsd[k] = sd[k]
        outs = oa.forward_training(wsd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                   batch["pitches"], batch["energies"], batch["stress_indices"])
        want = oa.training_losses(ocfg, outs, batch["mel_specs"], batch["phoneme_durations"],
                                  batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                  batch["mel_lengths"], batch["phoneme_lengths"])
        want = torch.tensor([float(x) for x in want])
        assert torch.allclose(g, want, rtol=1.5e-2, atol=1e-4), (use_ema, g, want)
    ts.train_step(batch)                                                # training continues normally afterwards
    assert int(ts.engine.drop_state[1]) == rng_step + 1


@pytest.mark.parametrize("graphs", [False, True])
def test_train_step_host_returns_each_steps_losses_early(graphs):
    """train_step_host() hands back the six losses of THE step it ran (sequence id check) — the numbers train_step()
    leaves in its device tensor — for varying batches and shapes, and trains the same weights."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    ocfg, cfg = _tiny()
    sd = oa.seeded_state_dict(ocfg, seed=0)
    mk = lambda: TrainStep(cfg, OptimConfig(learning_rate=1e-3), ScheduleConfig(total_steps=1000), device="cuda", use_graphs=graphs)
    a, b = mk(), mk()
    a.load_state_dict(sd)
    b.load_state_dict(sd)
    batches = [{k: v.pin_memory() for k, v in oa.synthetic_batch(B=3, P=24, T=T, seed=s, ragged=True).items()}
               for T, s in ((150, 11), (150, 12), (97, 13), (150, 11), (97, 14), (150, 15), (600, 16), (600, 17), (600, 16))]
    # (same shape, different mels back to back: the next mel is copied from a side stream while the step still runs)
    for i, batch in enumerate(batches):
        want = a.train_step(batch).cpu().tolist()
        got = b.train_step_host(batch)
        if i == 0:
            assert got == want, (got, want)       # same weights, deterministic forward: the very same numbers
        # later steps: the split-K weight gradients add with fp32 atomics, so two runs differ in the last bits
        assert all(abs(g - w) <= 2e-3 * abs(w) + 1e-5 for g, w in zip(got, want)), (i, got, want)
    torch.cuda.synchronize()
    num = den = 0.0
    mine, other = a.store.state_dict(), b.store.state_dict()
    for n in a.store.order:
        num += float((mine[n].float() - other[n].float()).pow(2).sum())
        den += float((mine[n].float().cpu() - sd[n]).pow(2).sum())
    assert (num / den) ** 0.5 < 0.15, (num / den) ** 0.5          # same gate as the oracle comparison above
