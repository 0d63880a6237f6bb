"""kokoro-train CLI surface and epoch loop, without a GPU: the parser is compared option by option with a fixture
dumped from the live reference parser (tests/golden/make_golden_cli.py); the loop runs on a stub step object."""
import argparse
import json
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _rows(parser):
    rows = []
    for a in parser._actions:
        if isinstance(a, argparse._HelpAction):
            continue
        rows.append({"options": sorted(a.option_strings), "dest": a.dest, "default": a.default,
                     "type": getattr(a.type, "__name__", None), "action": type(a).__name__})
    return rows


def test_parser_has_every_reference_flag_with_the_same_default():
    from kokoro_ruslan_b200.cli import build_parser
    want = json.load(open(os.path.join(HERE, "golden", "cli_flags.json")))
    mine = {tuple(r["options"]): r for r in _rows(build_parser())}
    for r in want:
        got = mine.get(tuple(r["options"]))
        assert got is not None, f"missing reference flag {r['options']}"
        assert (got["dest"], got["default"], got["type"], got["action"]) == (r["dest"], r["default"], r["type"], r["action"]), (got, r)
    extra = {k for k in mine if list(k) not in [r["options"] for r in want]}
    assert extra == {("--synthetic",), ("--seed",)}, extra


def test_config_mapping_follows_the_reference_rules():
    from kokoro_ruslan_b200.cli import build_parser, create_config_from_args
    cfg = create_config_from_args(build_parser().parse_args([]))
    assert (cfg.num_epochs, cfg.learning_rate, cfg.max_frames_per_batch, cfg.validation_split) == (30, 5e-5, 30000, 0.1)
    assert cfg.gradient_accumulation_steps == 2 and cfg.use_dynamic_batching
    cfg = create_config_from_args(build_parser().parse_args(
        ["-c", "corp", "-o", "out", "-b", "1", "-e", "1", "-lr", "1e-4", "--no-validation", "--no-dynamic-batching",
         "--max-frames", "8000", "--no-mfa"]))
    assert (cfg.data_dir, cfg.output_dir, cfg.batch_size, cfg.num_epochs, cfg.learning_rate) == ("corp", "out", 1, 1, 1e-4)
    assert cfg.validation_split == 0.0 and not cfg.use_dynamic_batching and cfg.max_frames_per_batch == 8000


def test_accumulation_windows_use_the_exact_tail_divisor():
    from kokoro_ruslan_b200.cli import accumulation_windows
    assert accumulation_windows(5, 2) == [[0, 1], [2, 3], [4]]
    assert accumulation_windows(4, 2) == [[0, 1], [2, 3]]
    assert accumulation_windows(3, 1) == [[0], [1], [2]]
    assert accumulation_windows(0, 2) == []


class _StubEngine:
    D = 512

    def __init__(self):
        self.spec_calls = []

    def set_spec_augment(self, spans, n_time=1, n_feat=2):
        self.spec_calls.append(None if spans is None else tuple(spans.shape))


class _StubStep:
    """TrainStep surface used by cli.train()."""

    def __init__(self, val_curve):
        self.engine = _StubEngine()
        self.calls, self.evals, self.val_curve = [], 0, list(val_curve)
        self.epoch_marker = 0
        self.sched = type("S", (), {"current_optimizer_step": 0, "state_dict": lambda s: {"k": 1}})()
        self.store = type("St", (), {"ema": None})()

    def micro_step(self, batch, first=True, last=True, divisor=1):
        self.calls.append((tuple(batch["mel_specs"].shape), first, last, divisor))
        if last:
            self.sched.current_optimizer_step += 1
        return torch.tensor([2.0, 1.0, 0.5, 0.25, 0.1, 0.1])

    def eval_losses(self, batch, use_ema=True):
        v = self.val_curve[min(self.epoch_marker, len(self.val_curve) - 1)]
        self.evals += 1
        return torch.tensor([v, v, 0.0, 0.0, 0.0, 0.0])

    def state_dict(self):
        return {"w": torch.zeros(2)}


def test_epoch_loop_windows_spec_augment_validation_and_early_stopping(tmp_path):
    from kokoro_ruslan_b200 import cli
    ds = cli.SyntheticDataset(24, seed=1, min_frames=60, max_frames=120)
    train_ds, val_ds = cli.split_dataset(ds, 0.25, seed=3)
    assert len(train_ds) == 18 and len(val_ds) == 6
    cfg = cli.RunConfig(output_dir=str(tmp_path), num_epochs=10, max_frames_per_batch=400, min_batch_size=1,
                        max_batch_size=4, early_stopping_patience=2, save_every=2)
    step = _StubStep([1.0, 0.9, 0.95, 0.96, 0.97])
    orig = step.eval_losses

    logs = []

    def log(s):
        logs.append(s)
        step.epoch_marker += 1
    out = cli.train(cfg, train_ds, val_ds, step, log=log)
    # val improves in epochs 1-2, then stalls: stop after 2 + patience epochs
    assert len(out["history"]) == 4 and out["best_val_epoch"] == 1 and abs(out["best_val_loss"] - 0.9) < 1e-6
    # every window starts with first=True, ends with last=True and carries its own length as divisor
    i = 0
    per_epoch = [h["batches"] for h in out["history"]]
    for nb in per_epoch:
        for win in cli.accumulation_windows(nb, 2):
            for k, _ in enumerate(win):
                shape, first, last, div = step.calls[i]
                assert (first, last, div) == (k == 0, k == len(win) - 1, len(win))
                assert shape[0] * shape[1] <= 400 or shape[0] == 1
                i += 1
    assert i == len(step.calls)
    # SpecAugment off in epoch 0, on afterwards (trainer.py:2042-2055)
    n0 = per_epoch[0]
    assert all(c is None for c in step.engine.spec_calls[:n0]) and all(c is not None for c in step.engine.spec_calls[n0:])
    # checkpoints: improvements (epochs 1, 2) + save_every (2, 4), reference file names and keys
    names = sorted(os.path.basename(p) for p in set(out["checkpoints"]))
    assert names == ["checkpoint_epoch_1.pth", "checkpoint_epoch_2.pth", "checkpoint_epoch_4.pth"], names
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint_epoch_2.pth"), weights_only=False)
    for key in ("epoch", "model_state_dict", "ema_model_state_dict", "current_optimizer_step", "optimizer_steps_completed",
                "scheduler_state_dict", "loss", "train_loss", "val_loss", "best_val_loss", "best_val_epoch", "config"):
        assert key in ck, key
    assert orig is not None


def test_main_refuses_to_run_without_cuda():
    from kokoro_ruslan_b200 import cli
    if torch.cuda.is_available():
        return
    try:
        cli.main(["--synthetic", "4", "-e", "1"])
    except RuntimeError as exc:
        assert "no CPU" in str(exc)
    else:
        raise AssertionError("expected a RuntimeError on a CPU-only host")


def test_resume_auto_picks_the_latest_checkpoint_and_continues(tmp_path):
    from kokoro_ruslan_b200 import cli
    for n in (1, 3, 12):
        torch.save({"epoch": n - 1, "model_state_dict": {"w": torch.full((2,), float(n))}, "ema_model_state_dict": None,
                    "scheduler_state_dict": {"k": n}, "loss": 1.0 / n}, os.path.join(str(tmp_path), f"checkpoint_epoch_{n}.pth"))
    assert cli.find_latest_checkpoint(str(tmp_path)).endswith("checkpoint_epoch_12.pth")
    assert cli.find_latest_checkpoint(str(tmp_path / "missing")) is None
    step = _StubStep([1.0])
    loaded = {}
    step.load_state_dict = lambda sd: loaded.update(sd)
    cfg = cli.RunConfig(output_dir=str(tmp_path), resume_checkpoint="auto", num_epochs=14)
    start = cli.resume(cfg, step, log=lambda s: None)
    assert start == 12 and float(loaded["w"][0]) == 12.0
    ds = cli.SyntheticDataset(6, seed=1, min_frames=60, max_frames=90)
    cfg.min_batch_size, cfg.max_frames_per_batch, cfg.save_every = 1, 400, 0
    out = cli.train(cfg, ds, None, step, log=lambda s: None, start_epoch=start)
    assert [h["epoch"] for h in out["history"]] == [12, 13]


def test_data_parallel_ranks_split_every_epoch_evenly():
    from kokoro_ruslan_b200 import cli
    ds = cli.SyntheticDataset(40, seed=2, min_frames=60, max_frames=200)
    cfg = cli.RunConfig(output_dir="/tmp/_kr_cli_dp", num_epochs=2, max_frames_per_batch=600, min_batch_size=1,
                        max_batch_size=6, save_every=0, gradient_accumulation_steps=2)
    seen = []
    for rank in (0, 1):
        step = _StubStep([1.0])
        out = cli.train(cfg, ds, None, step, rank=rank, world=2, log=lambda s: None)
        seen.append((out, step.calls))
    assert [h["batches"] for h in seen[0][0]["history"]] == [h["batches"] for h in seen[1][0]["history"]]
    assert len(seen[0][1]) == len(seen[1][1]) > 0            # same number of micro-steps on both ranks (no hang)
    assert [c[1:] for c in seen[0][1]] == [c[1:] for c in seen[1][1]]      # same window structure -> same collectives


def test_validation_metrics_are_one_accumulator_per_epoch(tmp_path):
    """cfg.val_metrics: the loop hands ONE accumulator per validation epoch to every eval_losses call and reads it once
    (reference trainer.py:1868-1916 syncs four times per utterance instead)."""
    from kokoro_ruslan_b200 import cli

    class Step(_StubStep):
        def __init__(self):
            super().__init__([1.0, 0.9])
            self.accs, self.reads = [], 0

        def new_val_metrics(self):
            self.accs.append(torch.zeros(8))
            return self.accs[-1]

        def eval_losses(self, batch, use_ema=True, metrics=None):
            assert metrics is self.accs[-1]
            metrics[:4] += torch.tensor([0.5, 1.0, 0.25, 1.0])
            return super().eval_losses(batch, use_ema)

        def read_val_metrics(self, acc):
            self.reads += 1
            return {"val_spectral_convergence": float(acc[0] / acc[1]), "val_f0_rmse": float(acc[2] / acc[3])}

    ds = cli.SyntheticDataset(16, seed=2, min_frames=60, max_frames=120)
    train_ds, val_ds = cli.split_dataset(ds, 0.25, seed=3)
    cfg = cli.RunConfig(output_dir=str(tmp_path), num_epochs=2, max_frames_per_batch=400, min_batch_size=1,
                        max_batch_size=4, save_every=0, val_metrics=True)
    step = Step()
    out = cli.train(cfg, train_ds, val_ds, step, log=lambda s: None)
    assert len(step.accs) == 2 and step.reads == 2
    for h in out["history"]:
        assert h["val_spectral_convergence"] == 0.5 and h["val_f0_rmse"] == 0.25
    # default: the stub without the metrics surface / cfg.val_metrics False keeps the plain call
    step2 = _StubStep([1.0])
    cfg.val_metrics = False
    cli.train(cfg, train_ds, val_ds, step2, log=lambda s: None)
    assert step2.evals > 0


def test_main_wires_configs_into_the_step(tmp_path, monkeypatch, capsys):
    """cli.main() end to end on a stand-in TrainStep (no device): the optimizer gets the half-life EMA decay computed
    from the epoch's optimizer steps (reference trainer.py:808-822), the schedule its total step count, the dropout
    configuration the reference's training values; checkpoints of the run carry optimizer state and model metadata."""
    import math
    from kokoro_ruslan_b200 import cli, parallel, train_step
    from kokoro_ruslan_b200.optim import FusedAdamW
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    made = {}

    class FakeStep(_StubStep):
        def __init__(self, model_cfg, optim_cfg, sched_cfg, device=None, process_group=None, dropout=None):
            super().__init__([1.0, 0.9, 0.8])
            made.update(model_cfg=model_cfg, optim_cfg=optim_cfg, sched_cfg=sched_cfg, device=device, dropout=dropout)
            tiny = ModelConfig(vocab_size=model_cfg.vocab_size, hidden_dim=128, n_encoder_layers=1, n_heads=2,
                               encoder_ff_dim=256, n_decoder_layers=1, decoder_ff_dim=256, max_decoder_seq_len=400,
                               variance_filter_size=64)
            self.store = ParamStore(tiny, torch.device("cpu"), with_ema=False)
            self.store.init_default = lambda seed=0: None
            self.opt = FusedAdamW(self.store, optim_cfg)
            self.engine.cfg = tiny

        def state_dict(self):
            return self.store.state_dict()

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(parallel, "init_distributed", lambda: (0, 0, 1))
    monkeypatch.setattr(train_step, "TrainStep", FakeStep)
    rc = cli.main(["--synthetic", "12", "-e", "2", "-o", str(tmp_path), "--max-frames", "2400", "--min-batch-size", "1",
                   "--max-batch-size", "4", "--save-every", "1", "--val-split", "0.25", "--seed", "5"])
    assert rc == 0 and "done: 2 epochs" in capsys.readouterr().out
    total = made["sched_cfg"].total_steps
    assert total % 2 == 0 and total >= 2
    opt_steps = total // 2
    assert made["optim_cfg"].ema_decay == pytest.approx(max(0.9, min(math.exp(-math.log(2) / opt_steps), 0.9999)))
    assert made["optim_cfg"].learning_rate == 5.0e-5 and made["device"] == "cuda:0"
    d = made["dropout"]
    assert (d.encoder, d.decoder, d.decoder_input, d.variance, d.stochastic_depth) == (0.15, 0.20, 0.15, 0.1, 0.1)
    ck = torch.load(os.path.join(str(tmp_path), "checkpoint_epoch_2.pth"), weights_only=False)
    assert ck["model_metadata"]["architecture"]["hidden_dim"] == 128
    # `config` is a pickled kokoro.training.config.TrainingConfig instance like the reference's (trainer.py:1994-2031)
    assert type(ck["config"]).__module__ == "kokoro.training.config" and type(ck["config"]).__name__ == "TrainingConfig"
    for key in ("global_step", "ema_updates", "scheduler_config", "grad_explosion_state", "optimizer_steps_completed"):
        assert key in ck, key
    assert len(ck["optimizer_state_dict"]["param_groups"]) == 10 and ck["config"].ema_half_life_epochs == 1.0


def test_async_checkpoints_are_complete_when_train_returns(tmp_path):
    from kokoro_ruslan_b200 import cli
    ds = cli.SyntheticDataset(12, seed=1, min_frames=60, max_frames=120)
    train_ds, val_ds = cli.split_dataset(ds, 0.25, seed=3)
    cfg = cli.RunConfig(output_dir=str(tmp_path), num_epochs=3, max_frames_per_batch=400, min_batch_size=1,
                        max_batch_size=4, save_every=1, async_checkpoints=True)
    out = cli.train(cfg, train_ds, val_ds, _StubStep([1.0, 0.9, 0.8]), log=lambda s: None)
    names = sorted(os.path.basename(p) for p in set(out["checkpoints"]))
    assert names == ["checkpoint_epoch_1.pth", "checkpoint_epoch_2.pth", "checkpoint_epoch_3.pth"]
    for n in names:
        ck = torch.load(os.path.join(str(tmp_path), n), weights_only=False)
        assert "model_state_dict" in ck and not os.path.exists(os.path.join(str(tmp_path), n + ".tmp"))


def test_kokoro_shim_import_paths_and_console_script():
    """the reference's import paths and console script exist (shim/kokoro + pyproject.toml), resolve to the
    B200 implementation, and the checkpoint `config` pickles as kokoro.training.config.TrainingConfig.  Run in a fresh
    interpreter: this process may hold the REFERENCE's kokoro package (oracle harness)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, pickle; sys.path[:0] = [%r, %r]\n"
        "import kokoro.cli.training as t, kokoro.cli.cli as c, kokoro.training.config as cfg\n"
        "import kokoro.data.dataset as d, kokoro.utils.lengths as l\n"
        "from kokoro_ruslan_b200 import cli, data, lengths\n"
        "assert t.main is cli.main and c.build_parser is cli.build_parser\n"
        "assert cfg.TrainingConfig is cli.TrainingConfig and d.collate_fn is data.collate_fn\n"
        "assert d.DynamicFrameBatchSampler is data.DynamicFrameBatchSampler and l.length_regulate is lengths.length_regulate\n"
        "conf = cli.create_config_from_args(cli.build_parser().parse_args(['--epochs', '3']))\n"
        "back = pickle.loads(pickle.dumps(cli._picklable_config(conf)))\n"
        "assert type(back) is cfg.TrainingConfig and back.num_epochs == 3\n"
        "print('ok')\n" % (root, os.path.join(root, "shim")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]
    py = open(os.path.join(root, "pyproject.toml")).read()
    assert 'kokoro-train = "kokoro.cli.training:main"' in py
    # the model / vocoder modules import the CUDA-side packages lazily enough to be importable without a device
    code2 = ("import sys; sys.path[:0] = [%r, %r]\n"
             "import kokoro.model.model as m, kokoro.inference.hifigan_vocoder as h\n"
             "from kokoro_ruslan_b200.model import KokoroModel\n"
             "assert m.KokoroModel is KokoroModel and hasattr(h, 'HiFiGANGenerator') and hasattr(h, 'load_hifigan_model')\n"
             "print('ok')\n" % (root, os.path.join(root, "shim")))
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr[-2000:]


def test_schedule_rewind_undoes_skipped_steps():
    from kokoro_ruslan_b200.train_step import ScheduleConfig, WarmupOneCycle
    a = WarmupOneCycle(1e-3, [1.0, 0.5], ScheduleConfig(total_steps=50, warmup_steps=5))
    b = WarmupOneCycle(1e-3, [1.0, 0.5], ScheduleConfig(total_steps=50, warmup_steps=5))
    for _ in range(12):
        a.advance()
    for _ in range(9):
        b.advance()
    a.rewind(3)
    assert a.state_dict() == b.state_dict() and a.lrs() == b.lrs()


def test_sharded_checkpoint_merges_to_the_single_file_payload(tmp_path):
    """checkpoint.save_sharded / load_sharded: every rank of a data-parallel run writes its share of the
    (replicated) tensors; the merged payload equals what a single torch.save would have held, a missing shard is an error, a
    plain checkpoint passes through load_sharded unchanged, and the shares are balanced."""
    import torch
    from kokoro_ruslan_b200 import checkpoint as ck
    g = torch.Generator().manual_seed(0)
    payload = {"epoch": 3, "model_state_dict": {f"w{i}": torch.randn(10 + 7 * i, 4, generator=g) for i in range(9)},
               "ema_model_state_dict": {f"w{i}": torch.randn(10 + 7 * i, 4, generator=g) for i in range(9)},
               "optimizer_state_dict": {"state": {0: {"step": torch.tensor(5.0), "exp_avg": torch.randn(33, generator=g)}},
                                        "param_groups": [{"lr": 1e-4, "params": [0], "betas": (0.9, 0.98)}]},
               "config": {"a": 1}, "best_val_loss": 0.5, "scheduler_state_dict": {"step": 7}, "none": None}
    world = 4
    path = str(tmp_path / "checkpoint_epoch_4.pth")
    written = []
    for r in range(world):                                   # what the four ranks do, one after the other here
        written += ck.save_sharded(path, payload, r, world)
    assert sorted(os.path.basename(f) for f in written) == sorted(
        ["checkpoint_epoch_4.pth"] + [f"checkpoint_epoch_4.pth.shard{r}of4" for r in range(world)])
    merged = ck.load_sharded(path)

    def same(a, b):
        if isinstance(a, torch.Tensor):
            return isinstance(b, torch.Tensor) and torch.equal(a, b)
        if isinstance(a, dict):
            return isinstance(b, dict) and list(a) == list(b) and all(same(a[k], b[k]) for k in a)
        if isinstance(a, (list, tuple)):
            return type(a) is type(b) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
        return a == b
    assert same(payload, merged)
    table = ck.shard_assignment(payload, world)
    load = [0] * world
    for p_, t in ck._tensor_paths(payload):
        load[table[p_]] += t.numel()
    assert max(load) - min(load) <= max(t.numel() for _, t in ck._tensor_paths(payload))
    head = torch.load(path, weights_only=False)
    assert ck.is_sharded(head) and not list(ck._tensor_paths(head))       # the head file holds no tensor
    os.remove(ck.shard_path(path, 2, world))
    with pytest.raises(FileNotFoundError):
        ck.load_sharded(path)
    plain = str(tmp_path / "plain.pth")
    torch.save(payload, plain)
    assert same(payload, ck.load_sharded(plain))


def test_cached_corpus_trains_through_the_prefetcher(tmp_path, monkeypatch):
    """A corpus whose features sit in the reference's cache format (N4) is trained on without the reference package: the
    CLI's dataset loader falls back to <corpus>/.feature_cache, and the epoch loop takes its batches from the read-ahead
    prefetcher — the same batches, in the same order, as the synchronous loop."""
    from kokoro_ruslan_b200 import cli
    from kokoro_ruslan_b200.data import collate_fn
    from kokoro_ruslan_b200.feature_cache import BatchPrefetcher, FeatureCache
    corpus = tmp_path / "corpus"
    cache = FeatureCache(corpus / ".feature_cache")
    src = cli.SyntheticDataset(14, seed=5, min_frames=50, max_frames=110)
    for i in range(len(src)):
        it = dict(src[i])
        it["mel_spec"] = it["mel_spec"] if "mel_spec" in it else it["mel"]
        cache.save(f"utt{i:02d}", {k: v for k, v in it.items() if k in ("mel_spec", "phoneme_indices", "stress_indices",
                                                                        "phoneme_durations", "stop_token_targets", "pitch",
                                                                        "energy", "text", "mel_length", "phoneme_length")})
    for name in ("kokoro", "kokoro.data", "kokoro.data.dataset", "kokoro.training", "kokoro.training.config"):
        monkeypatch.setitem(sys.modules, name, None)         # the reference package is not importable
    ds = cli.load_reference_dataset(cli.RunConfig(data_dir=str(corpus), output_dir=str(tmp_path)))
    assert len(ds) == 14 and ds.vocab_size >= 59 and ds.thread_safe
    train_ds, val_ds = cli.split_dataset(ds, 0.2, seed=1)
    assert train_ds.thread_safe
    cfg = cli.RunConfig(output_dir=str(tmp_path / "out"), num_epochs=2, max_frames_per_batch=300, min_batch_size=1, max_batch_size=4,
                        save_every=0)
    step = _StubStep([1.0, 0.9])
    out = cli.train(cfg, train_ds, val_ds, step)
    assert len(out["history"]) == 2 and sum(h["batches"] for h in out["history"]) == len(step.calls)
    # the prefetcher yields exactly the synchronous batches
    batches = [[0, 3], [1], [2, 4, 5]]
    want = [collate_fn([train_ds[i] for i in b], pin_memory=False) for b in batches]
    got = list(BatchPrefetcher(train_ds, batches, lambda items: collate_fn(items, pin_memory=False), depth=2, workers=2))
    assert len(got) == 3
    for a, b in zip(want, got):
        assert sorted(a) == sorted(b) and all(torch.equal(a[k], b[k]) for k in a if isinstance(a[k], torch.Tensor))
    with pytest.raises(ValueError):
        BatchPrefetcher(cli.SyntheticDataset(2), [[0]], collate_fn)
    empty = tmp_path / "nothing"
    empty.mkdir()
    with pytest.raises(RuntimeError):
        cli.load_reference_dataset(cli.RunConfig(data_dir=str(empty), output_dir=str(tmp_path)))


def test_legacy_checkpoint_migrations_follow_the_reference_rules():
    """checkpoint.migrate_model_state_dict / extract_model_state_dict / check_resume_fields restate the reference's
    strict-resume rules (training/checkpoint_manager.py:360-525): only variance-adaptor and ffn output-norm keys may be
    missing (they keep their initial values), only ALiBi buffers may be in excess (dropped); `model` and raw state dicts are
    accepted as the model entry; optimizer / scheduler / epoch / loss are required for a training resume."""
    from kokoro_ruslan_b200 import checkpoint as ck
    cur = {"encoder.w": torch.zeros(3), "duration_adaptor.variance_adaptor.pitch_embedding.weight": torch.ones(4),
           "decoder.layers.0.ff.output_norm.weight": torch.ones(2), "decoder.layers.0.ff.linear1.weight": torch.zeros(2, 2)}
    legacy = {"encoder.w": torch.full((3,), 5.0), "decoder.layers.0.ff.linear1.weight": torch.full((2, 2), 7.0),
              "decoder.layers.0.self_attn.alibi_slopes": torch.arange(8.0)}
    logs = []
    out = ck.migrate_model_state_dict(legacy, cur, logs.append)
    assert list(out) == list(cur) and torch.equal(out["encoder.w"], legacy["encoder.w"])
    assert torch.equal(out["duration_adaptor.variance_adaptor.pitch_embedding.weight"], torch.ones(4))
    assert torch.equal(out["decoder.layers.0.ff.output_norm.weight"], torch.ones(2))
    assert len(logs) == 2 and "alibi_slopes" in logs[1] and "missing 2 key(s)" in logs[0]
    assert ck.migrate_model_state_dict(dict(cur), cur) .keys() == cur.keys()
    for bad in (dict(legacy, **{"something.else": torch.zeros(1)}),                 # unexpected non-ALiBi key
                {k: v for k, v in legacy.items() if k != "encoder.w"},              # a missing key outside the migrations
                dict(legacy, **{"encoder.w": torch.zeros(4)})):                     # shape mismatch
        with pytest.raises(RuntimeError, match="architecture/state mismatch"):
            ck.migrate_model_state_dict(bad, cur)
    assert ck.extract_model_state_dict({"model_state_dict": legacy}) is legacy
    assert ck.extract_model_state_dict({"model": legacy}) is legacy
    assert ck.extract_model_state_dict(legacy) is legacy                              # raw state dict
    with pytest.raises(RuntimeError, match="recognized model state"):
        ck.extract_model_state_dict({"epoch": 3, "weights": legacy})
    with pytest.raises(RuntimeError, match="not a dictionary"):
        ck.extract_model_state_dict([1, 2])
    full = {"optimizer_state_dict": {}, "scheduler_state_dict": {}, "epoch": 1, "loss": 0.5}
    ck.check_resume_fields(full)
    with pytest.raises(RuntimeError, match="optimizer/scheduler"):
        ck.check_resume_fields({"epoch": 1, "loss": 0.5})
    with pytest.raises(RuntimeError, match="'epoch' or 'loss'"):
        ck.check_resume_fields({"optimizer_state_dict": {}, "scheduler_state_dict": {}, "epoch": 1})
    ck.check_resume_fields({"epoch": 1, "loss": 0.5}, training=False)


def test_phoneme_processor_pickle_is_read_by_the_references_loader(tmp_path):
    """cli.save_phoneme_processor writes what the reference trainer writes at the start of training (trainer.py:2828): the
    installed reference's load_phoneme_processor reads it back into an equal processor; datasets without a processor (the
    synthetic one) write nothing; the train / validation _Subset wrappers are looked through."""
    import logging
    sys.path.insert(0, os.path.dirname(HERE))
    from kokoro_ruslan_b200 import cli
    assert cli.save_phoneme_processor(cli.SyntheticDataset(2), str(tmp_path), log=lambda s: None) is None
    assert not (tmp_path / "phoneme_processor.pkl").exists()
    from oracle import ref_trainer as harness
    if not harness.reference_available():
        pytest.skip("baseline/_ref is not installed")
    harness._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.data.russian_phoneme_processor import RussianPhonemeProcessor
    from kokoro.training.checkpoint_manager import load_phoneme_processor
    proc = RussianPhonemeProcessor()
    corpus = type("Corpus", (), {"phoneme_processor": proc, "samples": [{"audio_length": 10}] * 3,
                                 "__len__": lambda s: 3, "__getitem__": lambda s, i: {}})()
    wrapped = cli._Subset(corpus, [0, 2])
    path = cli.save_phoneme_processor(wrapped, str(tmp_path), log=lambda s: None)
    assert path == str(tmp_path / "phoneme_processor.pkl")
    back = load_phoneme_processor(str(tmp_path))
    assert back.phoneme_to_id == proc.phoneme_to_id and back.get_vocab_size() == proc.get_vocab_size()
    norm = lambda d: {k: (sorted(v) if isinstance(v, list) else v) for k, v in d.items()}      # noqa: E731 — sets travel as lists
    assert norm(back.to_dict()) == norm(proc.to_dict())


def test_the_references_inference_loads_the_final_model_written_here(tmp_path):
    """The output directory of cli.train is what the reference's inference expects: KokoroTTS._load_phoneme_processor and
    KokoroTTS._load_model (inference/inference.py:85-330, unmodified, installed in baseline/_ref) read phoneme_processor.pkl
    and kokoro_russian_final.pth written by cli.save_phoneme_processor / cli.save_final_model, build the reference model from
    the metadata and end up with the very same 135 tensors."""
    import logging
    from pathlib import Path
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_trainer as harness
    if not harness.reference_available():
        pytest.skip("baseline/_ref is not installed")
    harness._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.data.russian_phoneme_processor import RussianPhonemeProcessor
    from kokoro.inference.inference import KokoroTTS
    from kokoro_ruslan_b200 import cli
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    proc = RussianPhonemeProcessor()
    mc = ModelConfig(vocab_size=len(proc.phoneme_to_id), mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2,
                     encoder_ff_dim=256, n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=400,
                     variance_filter_size=64, n_variance_bins=256)
    store = ParamStore(mc, torch.device("cpu"), with_ema=False)
    store.params.copy_(torch.randn(store.total, generator=torch.Generator().manual_seed(0)) * 0.05)
    step = type("Step", (), {"store": store, "engine": type("E", (), {"cfg": mc})(), "state_dict": lambda s: store.state_dict()})()
    cfg = cli.RunConfig(output_dir=str(tmp_path))
    assert cli.save_final_model(cfg, step) == str(tmp_path / "kokoro_russian_final.pth")
    cli.save_phoneme_processor(type("Corpus", (), {"phoneme_processor": proc})(), str(tmp_path), log=lambda s: None)
    tts = KokoroTTS.__new__(KokoroTTS)
    tts.model_dir, tts.device, tts.weights_preference, tts.enable_profiling = Path(tmp_path), torch.device("cpu"), "auto", False
    for k in ("inference_max_len", "inference_stop_threshold", "inference_min_len_ratio", "inference_min_len_floor"):
        setattr(tts, k, None)
        setattr(tts, "_explicit_" + k, False)
    tts.phoneme_processor = tts._load_phoneme_processor()
    assert tts.phoneme_processor.phoneme_to_id == proc.phoneme_to_id
    model = tts._load_model()
    mine, theirs = store.state_dict(), model.state_dict()
    assert len(mine) == 135 and set(mine) <= set(theirs)
    for k, v in mine.items():
        assert torch.equal(theirs[k], v), k
    assert tts.inference_max_len is not None and tts.inference_stop_threshold is not None     # controls came from the metadata
