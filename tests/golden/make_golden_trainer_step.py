"""Generates tests/golden/trainer_step.npz by driving the LIVE reference trainer's own methods
(/root/reference/src/kokoro/training/trainer.py) over seeded gradients — the pin of oracle/train_step.py's optimizer part
and of the product's FusedAdamW (SURVEY.md rows A12, A13, A13', A14).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden_trainer_step.py

A ``KokoroTrainer`` is built with ``__new__`` + the attributes its step methods read (the pattern of the reference's own
tests/unit/test_trainer_adaptive_stabilization.py:41-137), around a real (small) reference ``KokoroModel`` and a real
reference ``TrainingConfig``.  What runs is the reference's code:

    _setup_optimizer()                    trainer.py:446-689   -> the live 10-group table (per-parameter lr, weight decay)
    _setup_grad_explosion_tracker()       :914-933
    _setup_weight_norm_constraints()      :845-881             -> named_modules() lookup of the 12 + 12 FFN matrices
    per step, the optimizer-step boundary of train_epoch (:2345-2470):
        _preclip_projection_spikes()      :1332-1407
        total norm over named_parameters  :2354-2361           (inline in train_epoch: restated here, 4 lines)
        _compute_grad_explosion_threshold :1308-1330
        exploding -> clip = min(clip, 0.3); detector EMA update   :2371-2399   (inline: restated here)
        _has_nonfinite_gradients -> skip  :2401-2456
        _optimizer_step_with_clipping()   :3317-3342 -> runtime_policies.py:14-87 (clip_grad_norm_, AdamW.step, _update_ema)
        optimizer_steps_completed += 1; _apply_weight_norm_constraints()   :2466-2468, :883-912

Weights come from oracle.acoustic.seeded_state_dict, gradients from a seeded generator (``make_grads`` below, shared with
the tests), so the fixture holds only results: per-parameter group hyper-parameters, and per step the detector outputs
plus per-tensor norms and 16 sampled elements of the weights and the EMA weights.
"""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import acoustic as oa  # noqa: E402

N_SAMPLES = 16
SMALL = oa.AcousticConfig(vocab_size=20, hidden_dim=64, n_heads=1, n_encoder_layers=2, n_decoder_layers=2, ff_dim=96,
                          variance_filter=32, n_bins=16, max_len=64)

# case name -> (config overrides, weight scaling {name fragment: factor}, per-step gradient scale, per-step special)
CASES = {
    # plain steps: a hot first step (pre-clips + global clip engage), then small ones (no clipping at all)
    "plain": dict(over=dict(learning_rate=1e-3, ema_decay=0.9), wscale={}, gscale=[2.0, 0.05, 0.05], special=[None] * 3),
    # encoder AND decoder FFN matrices above the projection ceiling (ceiling lowered so the seeded weights exceed it)
    "projection": dict(over=dict(learning_rate=1e-3, ema_decay=0.9, dec_ffn_max_weight_norm=1.5),
                       wscale={"transformer_encoder_layers.0.ff.linear1.weight": 3.0,
                               "transformer_encoder_layers.1.ff.linear2.weight": 3.0,
                               "decoder.layers.0.ff.linear2.weight": 3.0}, gscale=[0.05, 0.05], special=[None] * 2),
    # explosion detector: warm-up floor interpolation, then the EMA threshold takes over, a spike trips it (clip -> 0.3),
    # a non-finite step is skipped, and the step after it proceeds
    "explosion": dict(over=dict(learning_rate=1e-3, ema_decay=0.9, grad_explosion_warmup_steps=3,
                                grad_explosion_warmup_floor=40.0, grad_explosion_abs_floor=2.0,
                                grad_explosion_min_ema_steps=2, grad_explosion_multiplier=3.0),
                      wscale={}, gscale=[0.02, 0.02, 0.02, 1.0, 0.02, 0.02, 5.0],
                      special=[None, None, None, None, "nan", None, None]),
}


def make_grads(names, shapes, case: str, step: int, scale: float):
    """Seeded gradients, one generator per (case, step); shared by the fixture generator and the tests."""
    g = torch.Generator().manual_seed(1000 * (sorted(CASES).index(case) + 1) + step)
    return {n: torch.randn(shapes[n], generator=g) * scale for n in names}


def scaled_state_dict(case: str):
    sd = oa.seeded_state_dict(SMALL, seed=3)
    for frag, f in CASES[case]["wscale"].items():
        sd[frag] = sd[frag] * f
    return sd


def _ref_trainer(case: str):
    sys.path.insert(0, "/root/reference/src")
    from kokoro.model.model import KokoroModel
    from kokoro.training.config import TrainingConfig
    from kokoro.training.trainer import KokoroTrainer
    cfg = SMALL
    model = KokoroModel(vocab_size=cfg.vocab_size, mel_dim=cfg.mel_dim, hidden_dim=cfg.hidden_dim,
                        n_encoder_layers=cfg.n_encoder_layers, n_heads=cfg.n_heads, encoder_ff_dim=cfg.ff_dim,
                        encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0,
                        n_decoder_layers=cfg.n_decoder_layers, decoder_ff_dim=cfg.ff_dim, max_decoder_seq_len=cfg.max_len,
                        variance_filter_size=cfg.variance_filter, variance_dropout=0.0, n_variance_bins=cfg.n_bins,
                        pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=False,
                        qk_norm=True, ffn_output_norm=True)
    model.load_state_dict(scaled_state_dict(case), strict=True)
    tc = TrainingConfig()
    tc.n_encoder_layers, tc.n_decoder_layers = cfg.n_encoder_layers, cfg.n_decoder_layers
    tc.use_fused_adamw = False
    for k, v in CASES[case]["over"].items():
        assert hasattr(tc, k) or k in ("grad_explosion_abs_floor", "grad_explosion_multiplier"), k   # getattr-defaulted
        setattr(tc, k, v)
    tr = KokoroTrainer.__new__(KokoroTrainer)
    tr.config, tr.model = tc, model
    tr.device, tr.device_type = torch.device("cpu"), "cpu"
    tr.use_mixed_precision, tr.scaler = False, None
    tr.scheduler_per_batch = False                       # the LR schedule is pinned separately (test_host_cpu.py)
    tr._step_scheduler_with_warmup = lambda: None
    tr.mixed_precision_stats = {k: 0 for k in ("scale_updates", "scale_decreases", "overflow_count", "successful_steps",
                                               "skipped_steps")}
    tr.optimizer_steps_completed = 0
    tr.current_optimizer_step = 0
    tr.dataloader = [None] * 10                          # only len() is read (EMA half-life rule; ema_decay is explicit here)
    tr._setup_optimizer()
    tr._setup_ema()
    tr._setup_weight_norm_constraints()
    tr._setup_grad_explosion_tracker()
    return tr, tc


def run_case(case: str):
    spec = CASES[case]
    tr, tc = _ref_trainer(case)
    model = tr.model
    names = [n for n, _ in model.named_parameters()]
    shapes = {n: tuple(p.shape) for n, p in model.named_parameters()}
    by_id = {id(p): n for n, p in model.named_parameters()}
    fix = {"names": np.array(names)}
    # A13': the live group table
    lr_of, wd_of, gidx = {}, {}, {}
    for gi, pg in enumerate(tr.optimizer.param_groups):
        for p in pg["params"]:
            lr_of[by_id[id(p)]], wd_of[by_id[id(p)]], gidx[by_id[id(p)]] = pg["lr"], pg["weight_decay"], gi
    fix["group_lr"] = np.array([lr_of[n] for n in names], dtype=np.float64)
    fix["group_wd"] = np.array([wd_of[n] for n in names], dtype=np.float64)
    fix["group_index"] = np.array([gidx[n] for n in names], dtype=np.int64)
    fix["n_groups"] = np.int64(len(tr.optimizer.param_groups))
    fix["wn_projected"] = np.array([any(p is w for w in tr._dec_ff_weights + tr._enc_ff_weights)
                                    for _, p in model.named_parameters()])
    per_step = {k: [] for k in ("total_norm", "threshold", "exploding", "clip_used", "skipped", "preclipped",
                                "w_norms", "w_samples", "ema_norms", "ema_samples")}
    for step, (gs, special) in enumerate(zip(spec["gscale"], spec["special"])):
        grads = make_grads(names, shapes, case, step, gs)
        if special == "nan":
            grads["decoder.layers.1.ff.linear1.weight"][3, 5] = float("nan")
        for n, p in model.named_parameters():
            p.grad = grads[n].clone()
        # ---- the optimizer-step boundary of train_epoch, trainer.py:2345-2470 ----
        clipped = tr._preclip_projection_spikes()
        total = 0.0
        for n, p in model.named_parameters():                               # :2354-2361
            if p.grad is not None:
                total += p.grad.data.norm(2).item() ** 2
        total = total ** 0.5
        thr, _floor, _ready = tr._compute_grad_explosion_threshold()
        nonfinite = tr._has_nonfinite_gradients()
        exploding = total > thr                                             # :2369
        clip = float(tc.max_grad_norm)
        if exploding:
            clip = min(clip, 0.3)                                           # :2392
        if not nonfinite:                                                   # detector EMA, :2396-2401 — see the note below
            if tr.grad_explosion_norm_ema is None:
                tr.grad_explosion_norm_ema = total
            else:
                a = tr.grad_explosion_ema_alpha
                tr.grad_explosion_norm_ema = a * tr.grad_explosion_norm_ema + (1 - a) * total
            tr.grad_explosion_ema_steps += 1
        # NOTE (documented deviation): the reference folds a NaN / inf norm into the detector's EMA
        # before skipping the step, after which its threshold silently degrades to the floor for the rest of the run;
        # oracle and product leave the EMA untouched on a skipped step, and so does this driver.
        if nonfinite:                                                       # :2401-2456
            tr.optimizer.zero_grad(set_to_none=True)
        else:
            ok, _ = tr._optimizer_step_with_clipping(clip_norm=clip, step_scheduler=True, update_ema=True)
            assert ok
            tr.optimizer_steps_completed += 1                               # :2466-2468
            tr._apply_weight_norm_constraints()
        per_step["total_norm"].append(total if np.isfinite(total) else -1.0)
        per_step["threshold"].append(thr)
        per_step["exploding"].append(int(exploding and not nonfinite))
        per_step["clip_used"].append(clip)
        per_step["skipped"].append(int(nonfinite))
        per_step["preclipped"].append(np.array([n in clipped for n in names]))
        ema_sd = tr.ema_model.state_dict()
        for key, src in (("w", dict(model.named_parameters())), ("ema", ema_sd)):
            norms, samples = [], []
            for n in names:
                t = src[n].detach().reshape(-1)
                norms.append(float(t.double().norm()))
                samples.append(t[torch.linspace(0, t.numel() - 1, N_SAMPLES).long()].numpy())
            per_step[key + "_norms"].append(np.array(norms))
            per_step[key + "_samples"].append(np.stack(samples))
    for k, v in per_step.items():
        fix[k] = np.stack([np.asarray(x) for x in v])
    return fix, tr


def check_oracle(case: str, fix) -> float:
    """Cross-check oracle.train_step.CpuTrainStep against what was just generated (every element, not just samples)."""
    from oracle.train_step import CpuTrainStep, StepPolicy
    spec = CASES[case]
    over = spec["over"]
    pol = StepPolicy(**{k: v for k, v in over.items() if k.startswith("grad_explosion")})
    ts = CpuTrainStep(SMALL, scaled_state_dict(case), lr=over["learning_rate"], ema_decay=over["ema_decay"],
                      wn_max=over.get("dec_ffn_max_weight_norm", 95.0), policy=pol)
    names = list(fix["names"])
    shapes = {n: tuple(ts.sd[n].shape) for n in names}
    worst = 0.0
    for step, (gs, special) in enumerate(zip(spec["gscale"], spec["special"])):
        grads = make_grads(names, shapes, case, step, gs)
        if special == "nan":
            grads["decoder.layers.1.ff.linear1.weight"][3, 5] = float("nan")
        for n in names:
            ts.sd[n].grad = grads[n].clone()
        ts.optimizer_step()
        assert ts.last["exploding"] == fix["exploding"][step], (case, step)
        assert ts.last["skip"] == fix["skipped"][step]
        assert abs(ts.last["threshold"] - fix["threshold"][step]) <= 1e-9 * abs(fix["threshold"][step])
        assert ts.last["clip_used"] == fix["clip_used"][step]
        for i, n in enumerate(names):
            for key, src in (("w", ts.sd), ("ema", ts.ema)):
                t = src[n].detach().reshape(-1)
                s = t[torch.linspace(0, t.numel() - 1, N_SAMPLES).long()].numpy()
                want = fix[key + "_samples"][step][i]
                worst = max(worst, float(np.abs(s - want).max() / (np.abs(want).max() + 1e-12)))
                wn = fix[key + "_norms"][step][i]
                worst = max(worst, abs(float(t.double().norm()) - wn) / (wn + 1e-12))
    return worst


if __name__ == "__main__":
    logging.disable(logging.CRITICAL)
    out = {}
    for case in CASES:
        fix, tr = run_case(case)
        for k, v in fix.items():
            out[f"{case}/{k}"] = v
        print(f"{case}: groups {int(fix['n_groups'])}, projected tensors {int(fix['wn_projected'].sum())}, "
              f"exploding {fix['exploding'].tolist()}, skipped {fix['skipped'].tolist()}, "
              f"clip {fix['clip_used'].tolist()}, thresholds {[round(float(x), 3) for x in fix['threshold']]}, "
              f"total norms {[round(float(x), 3) for x in fix['total_norm']]}, "
              f"pre-clipped per step {fix['preclipped'].sum(axis=1).tolist()}, "
              f"tensors on the projection ceiling after the last step "
              f"{[str(n) for n, w, p_ in zip(fix['names'], fix['w_norms'][-1], fix['wn_projected']) if p_ and abs(w - CASES[case]['over'].get('dec_ffn_max_weight_norm', 95.0)) < 1e-4]}; "
              f"oracle worst rel diff {check_oracle(case, fix):.2e}")
    np.savez_compressed(os.path.join(HERE, "trainer_step.npz"), **out)
    print("wrote", os.path.join(HERE, "trainer_step.npz"), os.path.getsize(os.path.join(HERE, "trainer_step.npz")), "bytes")
