"""Optimizer-state checkpoints in torch.optim.AdamW's own format and the asynchronous writer —
pure tensor plumbing, exercised on a CPU ParamStore against torch.optim.AdamW itself."""
import os

import pytest
import torch


def _tiny():
    from kokoro_ruslan_b200.optim import FusedAdamW, OptimConfig
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    cfg = ModelConfig(vocab_size=59, mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                      n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64,
                      n_variance_bins=256)
    store = ParamStore(cfg, torch.device("cpu"), with_ema=True)
    g = torch.Generator().manual_seed(0)
    store.params.copy_(torch.randn(store.total, generator=g) * 0.05)
    return store, FusedAdamW(store, OptimConfig())


def _torch_adamw(store, opt):
    """torch.optim.AdamW over reference-shaped copies of the parameters, grouped like the reference trainer."""
    from kokoro_ruslan_b200.checkpoint import group_layout
    from kokoro_ruslan_b200.optim import group_hparams
    hp = group_hparams(opt.cfg)
    params = {n: torch.nn.Parameter(store.ref_view(store.params, n).clone().contiguous()) for n in store.order}
    groups = [{"params": [params[n] for n in names], "lr": opt.cfg.learning_rate * hp[g][0], "weight_decay": hp[g][1]}
              for g, names in enumerate(group_layout(store.order))]
    return torch.optim.AdamW(groups, betas=opt.cfg.adam_betas, eps=opt.cfg.adam_eps), params


def test_layout_is_the_references_ten_groups():
    from kokoro_ruslan_b200.checkpoint import GROUP_NAMES, group_layout
    from kokoro_ruslan_b200.params import ModelConfig, param_specs
    names = [n for n, _ in param_specs(ModelConfig(vocab_size=59))]
    layout = group_layout(names)
    assert len(GROUP_NAMES) == 10 and [len(g) for g in layout] == [94, 12, 52, 20, 48, 48, 12, 18, 2, 2]   # SURVEY A13'
    flat = [n for g in layout for n in g]
    assert sorted(flat) == sorted(names) and len(set(flat)) == 308


def test_state_dict_loads_into_torch_adamw_and_back():
    from kokoro_ruslan_b200.checkpoint import group_layout, load_optimizer_state_dict, optimizer_state_dict
    store, opt = _tiny()
    # (1) torch -> ours: two real AdamW steps, then restore the moments into the flat buffers
    topt, params = _torch_adamw(store, opt)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):
        for p in params.values():
            p.grad = torch.randn(p.shape, generator=g)
        topt.step()
    tsd = topt.state_dict()
    assert load_optimizer_state_dict(opt, tsd) is True
    flat_names = [n for grp in group_layout(store.order) for n in grp]
    for i, n in enumerate(flat_names):
        assert torch.equal(store.ref_view(store.exp_avg, n), tsd["state"][i]["exp_avg"]), n
        assert torch.equal(store.ref_view(store.exp_avg_sq, n), tsd["state"][i]["exp_avg_sq"]), n
    from kokoro_ruslan_b200.optim import CTRL_FIELDS
    assert int(opt.ctrl[CTRL_FIELDS.index("step")]) == 2
    conv = [n for n in store.order if n.endswith("conv_layers.0.weight")][0]
    assert store.ref_view(store.exp_avg, conv).shape[-1] == 3            # reference conv shape [C_out, C_in, 3]
    # (2) ours -> torch: torch.optim.AdamW validates and accepts the snapshot, tensors identical
    sd = optimizer_state_dict(opt)
    assert [len(pg["params"]) for pg in sd["param_groups"]] == [len(grp) for grp in group_layout(store.order)]
    assert [pg["params"][0] for pg in sd["param_groups"]][0] == 0 and sd["param_groups"][-1]["params"][-1] == len(flat_names) - 1
    topt2, params2 = _torch_adamw(store, opt)
    topt2.load_state_dict(sd)
    st2 = topt2.state_dict()["state"]
    for i in range(len(flat_names)):
        assert torch.equal(st2[i]["exp_avg"], tsd["state"][i]["exp_avg"]) and float(st2[i]["step"]) == 2.0
    # (3) a fresh optimizer (no step yet) has no per-parameter state, like torch
    store3, opt3 = _tiny()
    assert optimizer_state_dict(opt3)["state"] == {}


def test_mismatching_layout_restarts_the_moments():
    from kokoro_ruslan_b200.checkpoint import load_optimizer_state_dict, optimizer_state_dict
    store, opt = _tiny()
    store.exp_avg.fill_(1.0)
    store.exp_avg_sq.fill_(2.0)
    sd = optimizer_state_dict(opt, step=5)
    collapsed = {"state": sd["state"], "param_groups": [dict(sd["param_groups"][0],
                                                             params=[i for pg in sd["param_groups"] for i in pg["params"]])]}
    msgs = []
    assert load_optimizer_state_dict(opt, collapsed, msgs.append) is False      # the torch.compile single-group collapse
    assert float(store.exp_avg.abs().sum()) == 0.0 and float(store.exp_avg_sq.abs().sum()) == 0.0
    assert msgs and "restart" in msgs[0]
    assert load_optimizer_state_dict(opt, None) is False


def test_cli_checkpoint_round_trip_with_moments(tmp_path):
    from kokoro_ruslan_b200 import cli
    store, opt = _tiny()
    g = torch.Generator().manual_seed(2)
    store.exp_avg.copy_(torch.randn(store.total, generator=g))
    store.exp_avg_sq.copy_(torch.rand(store.total, generator=g))
    want_m, want_v = store.exp_avg.clone(), store.exp_avg_sq.clone()
    from kokoro_ruslan_b200.optim import CTRL_FIELDS
    opt.ctrl[CTRL_FIELDS.index("step")] = 11

    class Step:
        def __init__(self, store, opt):
            self.store, self.opt = store, opt
            self.sched = type("S", (), {"current_optimizer_step": 11, "state_dict": lambda s: {"current_optimizer_step": 11,
                                                                                                "sched_step": 11},
                                        "load_state_dict": lambda s, sd: None})()

        def state_dict(self):
            return self.store.state_dict()

        def load_state_dict(self, sd):
            for n in self.store.order:
                self.store.ref_view(self.store.params, n).copy_(sd[n])

    cfg = cli.RunConfig(output_dir=str(tmp_path))
    path = cli.save_checkpoint(cfg, Step(store, opt), 3, {"train_loss": 1.0}, 0.5, 2)
    ck = torch.load(path, weights_only=False)
    assert set(ck["optimizer_state_dict"]) == {"state", "param_groups"} and len(ck["optimizer_state_dict"]["param_groups"]) == 10
    store2, opt2 = _tiny()
    store2.ema = None
    cfg.resume_checkpoint = path
    assert cli.resume(cfg, Step(store2, opt2), log=lambda s: None) == 4
    # padding between 64-element aligned tensors is not part of any tensor: compare tensor by tensor
    for n in store.order:
        assert torch.equal(store2.ref_view(store2.exp_avg, n), store.ref_view(want_m, n))
        assert torch.equal(store2.ref_view(store2.exp_avg_sq, n), store.ref_view(want_v, n))
    assert int(opt2.ctrl[CTRL_FIELDS.index("step")]) == 11


def test_async_writer_snapshots_and_surfaces_errors(tmp_path):
    from kokoro_ruslan_b200.checkpoint import AsyncCheckpointWriter
    w = AsyncCheckpointWriter()
    t = torch.arange(10.0)
    path = os.path.join(str(tmp_path), "a.pth")
    w.save(path, {"x": t, "nested": {"y": [t * 2, 3]}, "epoch": 4})
    t.add_(100.0)                                    # mutate after save(): the snapshot must not see it
    w.wait()
    back = torch.load(path, weights_only=False)
    assert torch.equal(back["x"], torch.arange(10.0)) and torch.equal(back["nested"]["y"][0], torch.arange(10.0) * 2)
    assert back["epoch"] == 4 and not os.path.exists(path + ".tmp")
    w.save(os.path.join(str(tmp_path), "no_such_dir", "b.pth"), {"x": t})
    with pytest.raises(RuntimeError):
        w.wait()


def test_model_metadata_matches_the_reference_schema_and_gates_resume(tmp_path):
    from kokoro_ruslan_b200 import cli
    from kokoro_ruslan_b200.checkpoint import build_model_metadata, check_model_metadata
    from kokoro_ruslan_b200.params import ModelConfig
    mc = ModelConfig(vocab_size=59)
    meta = build_model_metadata(mc, cli.RunConfig())
    # key set of build_model_metadata, reference training/checkpoint_manager.py:178-241
    assert meta["schema_version"] == 2
    assert set(meta["architecture"]) == {
        "mel_dim", "hidden_dim", "n_encoder_layers", "n_decoder_layers", "n_heads", "encoder_ff_dim", "decoder_ff_dim",
        "encoder_dropout", "max_decoder_seq_len", "use_variance_predictor", "variance_filter_size", "variance_kernel_size",
        "variance_dropout", "n_variance_bins", "pitch_min", "pitch_max", "energy_min", "energy_max", "use_stochastic_depth",
        "stochastic_depth_rate", "qk_norm", "ffn_output_norm", "vocab_size"}
    assert set(meta["inference_controls"]) == {"max_len", "stop_threshold", "min_len_ratio", "min_len_floor"}
    assert meta["architecture"]["encoder_ff_dim"] == 1536 and meta["architecture"]["encoder_dropout"] == 0.15
    assert check_model_metadata(meta, mc) == [] and check_model_metadata(None, mc) == []
    other = ModelConfig(vocab_size=59, hidden_dim=256, n_heads=4)
    bad = check_model_metadata(meta, other)
    assert len(bad) == 2 and bad[0].startswith("hidden_dim")
    meta["architecture"]["pitch_max"] = 800.0
    assert any("pitch_max" in b for b in check_model_metadata(meta, mc))

    # through cli.save_checkpoint / cli.resume
    store, opt = _tiny()
    store.ema = None                                  # (the EMA restore path refreshes the bf16 shadow: a device op)

    class Step:
        def __init__(self, cfg):
            self.store, self.opt = store, opt
            self.engine = type("E", (), {"cfg": cfg})()
            self.sched = type("S", (), {"current_optimizer_step": 0, "state_dict": lambda s: {}})()

        def state_dict(self):
            return self.store.state_dict()

        def load_state_dict(self, sd):
            pass

    tiny = ModelConfig(vocab_size=59, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256, n_decoder_layers=2,
                       decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64)
    cfg = cli.RunConfig(output_dir=str(tmp_path))
    path = cli.save_checkpoint(cfg, Step(tiny), 0, {}, 1.0, 0)
    assert torch.load(path, weights_only=False)["model_metadata"]["architecture"]["hidden_dim"] == 128
    cfg.resume_checkpoint = path
    assert cli.resume(cfg, Step(tiny), log=lambda s: None) == 1
    with pytest.raises(RuntimeError, match="different architecture"):
        cli.resume(cfg, Step(mc), log=lambda s: None)


def test_the_installed_references_own_loader_accepts_a_checkpoint_written_here(tmp_path):
    """Interop in the direction a reference user needs ("checkpoint interop"): a checkpoint written by
    cli.save_checkpoint goes through the UNMODIFIED reference's training/checkpoint_manager.py::load_checkpoint (strict
    resume: architecture metadata, strict model load, optimizer + scheduler state, epoch / loss) into the reference trainer's
    own model, 10-group AdamW and OneCycleLR — all 135 state-dict tensors and the Adam moments of all 132 parameters arrive
    bit-identical.  (The phoneme front-end is out of scope: its pickle is written by the reference's own helper.)"""
    import logging
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_trainer as harness
    if not harness.reference_available():
        pytest.skip("baseline/_ref is not installed")
    from kokoro_ruslan_b200 import cli
    from kokoro_ruslan_b200.optim import CTRL_FIELDS
    from kokoro_ruslan_b200.params import ModelConfig
    over = dict(n_mels=80, hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, encoder_ff_dim=256,
                decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64, n_variance_bins=256, num_epochs=3)
    tr = harness.build_trainer(None, torch.device("cpu"), 59, over)
    tr._setup_scheduler()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    store, opt = _tiny()
    g = torch.Generator().manual_seed(2)
    store.exp_avg.copy_(torch.randn(store.total, generator=g))
    store.exp_avg_sq.copy_(torch.rand(store.total, generator=g))
    opt.ctrl[CTRL_FIELDS.index("step")] = 11
    mc = ModelConfig(vocab_size=59, mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                     n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64, n_variance_bins=256)

    class Step:
        def __init__(self):
            self.store, self.opt = store, opt
            self.engine = type("E", (), {"cfg": mc})()
            self.sched = type("S", (), {"current_optimizer_step": 11,
                                        "state_dict": lambda s: {"current_optimizer_step": 11, "sched_step": 11},
                                        "load_state_dict": lambda s, sd: None})()

        def state_dict(self):
            return self.store.state_dict()

    path = cli.save_checkpoint(cli.RunConfig(output_dir=str(tmp_path)), Step(), 3, {"train_loss": 1.25}, 0.5, 2)
    from kokoro.data.russian_phoneme_processor import RussianPhonemeProcessor
    from kokoro.training.checkpoint_manager import load_checkpoint, save_phoneme_processor
    save_phoneme_processor(RussianPhonemeProcessor(), str(tmp_path))
    start_epoch, best_loss, _ = load_checkpoint(path, tr.model, tr.optimizer, tr.scheduler, str(tmp_path))
    assert start_epoch == 4 and best_loss == 1.25
    mine = store.state_dict()
    theirs = tr.model.state_dict()
    assert set(theirs) == set(mine) and len(mine) == 135
    for k, v in theirs.items():
        assert torch.equal(v, mine[k].cpu()), k
    names = {id(p): n for n, p in tr.model.named_parameters()}
    n_params = 0
    for grp in tr.optimizer.param_groups:
        for p in grp["params"]:
            n, st = names[id(p)], tr.optimizer.state[p]
            assert int(st["step"]) == 11, n
            assert torch.equal(st["exp_avg"], store.ref_view(store.exp_avg, n).cpu()), n
            assert torch.equal(st["exp_avg_sq"], store.ref_view(store.exp_avg_sq, n).cpu()), n
            n_params += 1
    assert n_params == 132 and len(tr.optimizer.param_groups) == 10


def test_resume_from_a_checkpoint_written_by_the_installed_reference_trainer(tmp_path):
    """The other direction: the UNMODIFIED reference trainer takes two AdamW steps and writes its own checkpoint
    (trainer.py:1994-2031, save_checkpoint_with_scaler: pickled TrainingConfig, model_metadata, OneCycleLR state, EMA);
    cli.resume loads it — weights, the Adam moments of every parameter and the step count arrive bit-identical, and training
    continues at the next epoch."""
    import logging
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import ref_trainer as harness
    if not harness.reference_available():
        pytest.skip("baseline/_ref is not installed")
    from kokoro_ruslan_b200 import cli
    from kokoro_ruslan_b200.optim import CTRL_FIELDS
    from kokoro_ruslan_b200.params import ModelConfig
    over = dict(n_mels=80, hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, encoder_ff_dim=256,
                decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64, n_variance_bins=256, num_epochs=3,
                output_dir=str(tmp_path))
    tr = harness.build_trainer(None, torch.device("cpu"), 59, over)
    tr._setup_scheduler()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    g = torch.Generator().manual_seed(5)
    for p in tr.model.parameters():
        p.grad = torch.randn(p.shape, generator=g) * 0.01
    tr.optimizer.step()
    tr.optimizer.step()
    tr.current_optimizer_step = tr.optimizer_steps_completed = tr.ema_updates = 2
    tr.save_checkpoint_with_scaler(2, 0.9, val_loss=1.1, best_val_loss=1.0, best_val_epoch=1)
    path = os.path.join(str(tmp_path), "checkpoint_epoch_3.pth")
    store, opt = _tiny()
    store.ema = None                                  # (the EMA restore path refreshes the bf16 shadow: a device op)
    mc = ModelConfig(vocab_size=59, mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                     n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64, n_variance_bins=256)

    class Step:
        def __init__(self):
            self.store, self.opt = store, opt
            self.engine = type("E", (), {"cfg": mc})()
            self.sched = type("S", (), {"current_optimizer_step": 0, "state_dict": lambda s: {},
                                        "load_state_dict": lambda s, sd: None})()

        def state_dict(self):
            return self.store.state_dict()

        def load_state_dict(self, sd):
            for n in self.store.order:
                self.store.ref_view(self.store.params, n).copy_(sd[n])

    cfg = cli.RunConfig(output_dir=str(tmp_path))
    cfg.resume_checkpoint = path
    assert cli.resume(cfg, Step(), log=lambda s: None) == 3
    theirs, mine = tr.model.state_dict(), store.state_dict()
    for k in mine:
        assert torch.equal(theirs[k], mine[k]), k
    names = {id(p): n for n, p in tr.model.named_parameters()}
    for grp in tr.optimizer.param_groups:
        for p in grp["params"]:
            n, st = names[id(p)], tr.optimizer.state[p]
            assert torch.equal(st["exp_avg"], store.ref_view(store.exp_avg, n)), n
            assert torch.equal(st["exp_avg_sq"], store.ref_view(store.exp_avg_sq, n)), n
    assert int(opt.ctrl[CTRL_FIELDS.index("step")]) == 2
