"""CPU-side checks: oracle pinned to the golden fixtures, parameter table, groups, C-ABI exports."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _cases():
    from oracle import acoustic as oa
    return {
        "tiny": (oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2,
                                   ff_dim=256, variance_filter=64, max_len=1200),
                 dict(B=3, P=24, T=150, seed=11, ragged=True)),
        "chunked": (oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1,
                                      ff_dim=128, variance_filter=64, max_len=1200),
                    dict(B=2, P=40, T=600, seed=12, ragged=True)),
        "full_width": (oa.AcousticConfig(max_len=1200), dict(B=2, P=32, T=200, seed=13, ragged=True)),
    }


@pytest.mark.parametrize("name", ["tiny", "chunked", "full_width"])
def test_oracle_matches_reference_golden(name):
    """The fp32 oracle reproduces the live reference's outputs, losses and gradients (fixtures made
    by tests/golden/make_golden.py from /root/reference)."""
    from oracle import acoustic as oa
    ocfg, bk = _cases()[name]
    batch = oa.synthetic_batch(n_mels=ocfg.mel_dim, vocab=ocfg.vocab_size, **bk)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    outs = oa.forward_training(sdr, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                               batch["pitches"], batch["energies"], batch["stress_indices"])
    losses = oa.training_losses(ocfg, outs, batch["mel_specs"], batch["phoneme_durations"],
                                batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                batch["mel_lengths"], batch["phoneme_lengths"])
    losses[0].backward()
    fix = np.load(os.path.join(HERE, "golden", f"acoustic_{name}.npz"))
    for key, got in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs):
        want = torch.from_numpy(fix[f"out_{key}"])
        assert torch.allclose(got.detach(), want, rtol=1e-4, atol=2e-5), key
    assert np.allclose([float(x) for x in losses], fix["losses"], rtol=1e-5, atol=1e-6)
    for n, norm, samp in zip(fix["grad_names"], fix["grad_norms"], fix["grad_samples"]):
        g = sdr[str(n)].grad
        if norm < 0:
            assert g is None or float(g.abs().max()) == 0.0
            continue
        assert abs(float(g.double().norm()) - norm) <= 1e-4 * norm + 1e-7, n
        flat = g.reshape(-1)
        idx = torch.linspace(0, flat.numel() - 1, 4).long()
        assert np.allclose(flat[idx].numpy(), samp, rtol=1e-3, atol=1e-6), n


def test_length_regulator_reference_vectors():
    """Exact expansion values the reference's own tests pin (tests/unit/test_utils_lengths.py:11-43)."""
    from oracle import acoustic as oa
    tokens = torch.tensor([[1.0, 2.0, 3.0]])
    out = oa.expand_tokens(tokens, torch.tensor([[2, 0, 3]]))
    assert out.tolist() == [[1.0, 1.0, 3.0, 3.0, 3.0]]
    out = oa.expand_tokens(tokens, torch.tensor([[1, 1, 1]]), max_len=5)
    assert out.tolist() == [[1.0, 2.0, 3.0, 0.0, 0.0]]
    out = oa.expand_tokens(tokens, torch.tensor([[3, 3, 3]]), max_len=4)
    assert out.tolist() == [[1.0, 1.0, 1.0, 2.0]]
    out = oa.expand_tokens(tokens, torch.tensor([[0, 0, 0]]))
    assert out.tolist() == [[0.0]]


def test_param_table_matches_reference_names():
    from kokoro_ruslan_b200.params import ModelConfig, param_specs
    fix = np.load(os.path.join(HERE, "golden", "acoustic_full_width.npz"))
    names = [n for n, _ in param_specs(ModelConfig())]
    assert names == [str(n) for n in fix["grad_names"]]
    assert len(names) == 308
    total = 0
    for _, shape in param_specs(ModelConfig()):
        n = 1
        for s in shape:
            n *= s
        total += n
    assert total == 49432276


def test_optimizer_groups_match_reference_counts():
    """SURVEY.md §8 A13': 94/12/52/20/48/48/12/18/2/2 tensors in the ten AdamW groups."""
    from kokoro_ruslan_b200.optim import OptimConfig, group_of, preclip_of
    from kokoro_ruslan_b200.params import ModelConfig, param_specs
    counts = [0] * 10
    elems = [0] * 10
    n_pre = 0
    for name, shape in param_specs(ModelConfig()):
        g = group_of(name)
        counts[g] += 1
        n = 1
        for s in shape:
            n *= s
        elems[g] += n
        n_pre += preclip_of(name, OptimConfig()) > 0
    assert counts == [94, 12, 52, 20, 48, 48, 12, 18, 2, 2]
    assert elems == [6365312, 14155776, 1785683, 91136, 12582912, 8448, 14155776, 24576, 262144, 513]
    assert n_pre == 126


def test_c_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "kokoro_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(kr_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) > 30
    so = os.path.join(ROOT, "kokoro_ruslan_b200", "libkokoro_b200.so")
    if not os.path.exists(so):
        from kokoro_ruslan_b200.build import build
        build()
    lib = ctypes.CDLL(so)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    lib.kr_abi_version.restype = ctypes.c_int
    assert lib.kr_abi_version() == 3


def test_ops_refuse_cpu_tensors():
    from kokoro_ruslan_b200 import ops
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, a, torch.zeros(128, 128))


def test_warmup_onecycle_matches_reference_stepping():
    """WarmupOneCycle == the reference's hand-rolled warm-up + torch OneCycleLR driven exactly as
    trainer.py:1519-1575 drives it (incl. the full-LR first step)."""
    import warnings
    from kokoro_ruslan_b200.train_step import ScheduleConfig, WarmupOneCycle
    total, warm, lr = 3000, 1200, 5e-5
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=lr)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        s = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=lr, total_steps=total - warm, pct_start=0.2,
                                                anneal_strategy="cos", cycle_momentum=False, div_factor=1.0,
                                                final_div_factor=1e4)
        w = WarmupOneCycle(lr, [1.0, 0.65], ScheduleConfig(total_steps=total))
        cur = 0
        for k in range(total + 20):
            ref = opt.param_groups[0]["lr"]
            got = w.lrs()
            assert abs(got[0] - ref) <= 1e-12 + 1e-6 * ref, (k, got, ref)
            assert abs(got[1] - 0.65 * ref) <= 1e-12 + 1e-6 * ref
            if cur < warm:
                opt.param_groups[0]["lr"] = lr * 0.01 + (lr - lr * 0.01) * cur / warm
            elif s.last_epoch < total - warm - 1:
                s.step()
            cur += 1
            w.advance()


def test_oracle_and_product_agree_on_groups_and_preclip():
    from kokoro_ruslan_b200.optim import OptimConfig, group_hparams, group_of, preclip_of
    from kokoro_ruslan_b200.params import ModelConfig, param_specs
    from oracle.train_step import GROUPS, param_group, spike_clip
    hp = group_hparams(OptimConfig())
    for (_, m, wd), (m2, wd2) in zip(GROUPS, hp):
        assert (m, wd) == (m2, wd2)
    for name, _ in param_specs(ModelConfig()):
        assert group_of(name) == param_group(name), name
        assert preclip_of(name, OptimConfig()) == spike_clip(name), name


def _trainer_fixture():
    import numpy as np
    return np.load(os.path.join(HERE, "golden", "trainer_step.npz"))


def test_group_table_preclip_and_projection_sets_match_the_live_trainer():
    """A13' / A12 / A14 against tests/golden/trainer_step.npz, generated by the LIVE KokoroTrainer's _setup_optimizer,
    _preclip_projection_spikes and _setup_weight_norm_constraints (tests/golden/make_golden_trainer_step.py)."""
    from kokoro_ruslan_b200.optim import OptimConfig, group_hparams, group_of, preclip_of, wnmax_of
    from oracle.train_step import GROUPS, param_group, spike_clip, wn_projected
    fx = _trainer_fixture()
    names = [str(n) for n in fx["plain/names"]]
    lr = 1e-3                                              # the fixture's base learning rate
    hp = group_hparams(OptimConfig(learning_rate=lr))
    assert int(fx["plain/n_groups"]) == 10
    for i, n in enumerate(names):
        m, wd = hp[group_of(n)]
        assert abs(lr * m - float(fx["plain/group_lr"][i])) < 1e-12 and wd == float(fx["plain/group_wd"][i]), n
        assert group_of(n) == int(fx["plain/group_index"][i]) == param_group(n), n     # live groups are in our order
        assert (wnmax_of(n, OptimConfig()) > 0) == bool(fx["plain/wn_projected"][i]) == wn_projected(n), n
    assert sum(bool(x) for x in fx["plain/wn_projected"]) == 8     # (2 encoder + 2 decoder layers) x linear1, linear2
    # pre-clip membership: in the hot first step of "plain" every tensor with a ceiling exceeds it
    for i, n in enumerate(names):
        has = preclip_of(n, OptimConfig()) > 0
        assert has == (spike_clip(n) > 0)
        if fx["plain/preclipped"][0][i]:
            assert has, n
    hot = {n for i, n in enumerate(names) if fx["explosion/preclipped"][6][i]}
    assert hot == {n for n in names if preclip_of(n, OptimConfig()) > 0}


def test_oracle_optimizer_step_matches_the_live_trainer():
    """oracle.train_step.CpuTrainStep.optimizer_step vs the live trainer: detector outputs exactly, weights and EMA
    weights to 1e-6 (measured 0: it is the same torch arithmetic in the same order)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden_trainer_step as mk
    fx = _trainer_fixture()
    for case in mk.CASES:
        fix = {k.split("/", 1)[1]: fx[k] for k in fx.files if k.startswith(case + "/")}
        assert mk.check_oracle(case, fix) < 1e-6, case
    assert fx["explosion/exploding"].tolist() == [0, 0, 0, 1, 0, 0, 1] and fx["explosion/skipped"].tolist() == [0, 0, 0, 0, 1, 0, 0]


def test_adaptive_stabilisation_values():
    from kokoro_ruslan_b200.train_step import adaptive_stabilisation
    assert adaptive_stabilisation(800, 20, 1.5) == (1.0, 1.5)
    s, c = adaptive_stabilisation(2000, 20, 1.5)          # 0.70 / 0.418
    assert abs(s - 0.7) < 1e-6 and abs(c - 0.5 / (2000 / 1400) ** 0.5) < 1e-9
    s, c = adaptive_stabilisation(800, 1200, 1.5)
    assert s == 0.25 and abs(c - 0.5 / 8 ** 0.5) < 1e-9


def test_only_tests_smoke_and_bench_import_the_oracle():
    """The oracle is test infrastructure: nothing in the product package or the developer tools may import it
    (tests/, __graft_entry__.smoke() and bench.py's CPU legs are the only users)."""
    import ast
    import glob
    root = os.path.dirname(HERE)
    offenders = []
    for path in glob.glob(os.path.join(root, "kokoro_ruslan_b200", "*.py")) + glob.glob(os.path.join(root, "tools", "*.py")):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            if any(m == "oracle" or m.startswith("oracle.") for m in mods):
                offenders.append(os.path.relpath(path, root))
    assert not offenders, offenders


def test_batch_validation_error_conventions():
    """KeyError / ValueError / TypeError as the reference's _transfer_batch_to_device raises them (trainer.py:1262-1297)."""
    import pytest
    import torch
    from kokoro_ruslan_b200.train_step import BATCH_KEYS, validate_batch
    from oracle import acoustic as oa
    batch = oa.synthetic_batch(B=2, P=8, T=20, seed=1)
    validate_batch(batch)
    for k in BATCH_KEYS:
        bad = {kk: v for kk, v in batch.items() if kk != k}
        with pytest.raises(KeyError):
            validate_batch(bad)
    with pytest.raises(ValueError):
        validate_batch({**batch, "pitches": None})
    with pytest.raises(TypeError):
        validate_batch({**batch, "mel_lengths": [20, 20]})
    with pytest.raises(ValueError):
        validate_batch({**batch, "energies": torch.zeros(3, 20)})


def test_recommended_ema_decay_matches_reference_formula():
    """utils/ema.py:6-27 as the trainer calls it (trainer.py:808-822: n_train = optimizer steps per epoch, batch_size 1)."""
    import math
    from kokoro_ruslan_b200.optim import recommended_ema_decay
    assert recommended_ema_decay(0, 1, 1.0) == 0.9999 and recommended_ema_decay(10, 0, 1.0) == 0.9999
    assert recommended_ema_decay(100, 1, 0.0) == 0.9999
    assert recommended_ema_decay(1000, 1, 1.0) == pytest.approx(math.exp(-math.log(2) / 1000))
    assert recommended_ema_decay(3, 1, 1.0) == 0.9                       # clipped below
    assert recommended_ema_decay(10 ** 6, 1, 1.0) == 0.9999              # clipped above
    assert recommended_ema_decay(5000, 10, 2.0) == pytest.approx(math.exp(-math.log(2) / 1000))


def test_oracle_explicit_padding_masks_match_reference_golden():
    """Explicit text_padding_mask / mel_padding_mask (model/model.py:586-589, 648-654): oracle vs the live-reference
    fixture tests/golden/acoustic_masks.npz (make_golden_masks.py)."""
    import sys
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from oracle import acoustic as oa
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    text = batch["phoneme_indices"] == 0
    for b in range(3):
        text[b, int(batch["phoneme_lengths"][b]) - 1] = True
    mel = torch.arange(150).unsqueeze(0) >= batch["mel_lengths"].unsqueeze(1)
    with torch.no_grad():
        outs = oa.forward_training(sd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                   batch["pitches"], batch["energies"], batch["stress_indices"], text_padding_mask=text,
                                   mel_padding_mask=mel)
    fix = np.load(os.path.join(HERE, "golden", "acoustic_masks.npz"))
    for k, o in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs):
        assert np.abs(o.numpy() - fix[f"out_{k}"]).max() < 1e-4, k


def test_param_views_are_cached_per_buffer_and_follow_repointed_buffers():
    """ParamStore.p / g / w hand out cached internal-layout views; the flat buffers get re-pointed (EMA weights during
    validation, symmetric-memory gradients under data parallelism) and the views must follow the buffer, not the name."""
    import torch
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    cfg = ModelConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1, encoder_ff_dim=256,
                      decoder_ff_dim=256, variance_filter_size=64)
    st = ParamStore(cfg, torch.device("cpu"), with_ema=True)
    name = "transformer_encoder_layers.0.ff.linear1.weight"
    conv = next(n for n in st.order if ".conv_layers.0.weight" in n)
    assert st.p(name) is st.p(name) and st.g(name) is st.g(name) and st.w(name) is st.w(name)
    assert st.p(name).data_ptr() == st.params.data_ptr() + 4 * st.entries[name].offset
    co, ci, k = st.entries[conv].shape
    assert tuple(st.p(conv).shape) == (co, k * ci)                      # internal conv layout [C_out, 3 * C_in]
    old = st.g(name)
    st.grads = torch.zeros_like(st.grads)                                # what parallel.SymmetricGradReducer does
    new = st.g(name)
    assert new is not old and new.data_ptr() == st.grads.data_ptr() + 4 * st.entries[name].offset
    saved = st.params
    st.params = st.ema                                                   # what TrainStep.eval_losses does
    assert st.p(name).data_ptr() == st.ema.data_ptr() + 4 * st.entries[name].offset
    st.params = saved
    assert st.p(name).data_ptr() == saved.data_ptr() + 4 * st.entries[name].offset
    st.p(name).fill_(3.0)
    assert float(st.params[st.entries[name].offset]) == 3.0 and float(st.ema[st.entries[name].offset]) == 0.0


def test_graph_capture_follows_the_shape_cache_hit_rate():
    """Fixed shapes are captured at once; after the probe window capturing goes on only while shapes keep coming back
    (dynamic batching produces a new shape almost every step: tools/dynamic_bench.py)."""
    from kokoro_ruslan_b200.train_step import TrainStep
    ts = TrainStep.__new__(TrainStep)
    probe = TrainStep.CAPTURE_PROBE
    ts._tick, ts._shape_hits = 2, 1                       # second step of a fixed-shape run
    assert ts._capture_worth()
    ts._tick, ts._shape_hits = probe, 0                   # still probing: a recurring shape is captured
    assert ts._capture_worth()
    ts._tick, ts._shape_hits = probe + 1, 1               # dynamic batching: 1 hit in 33 steps
    assert not ts._capture_worth()
    ts._tick, ts._shape_hits = 10 * probe, 9 * probe      # fixed shape, long run
    assert ts._capture_worth()
    ts._tick, ts._shape_hits = 1000, 499
    assert not ts._capture_worth()
    ts._tick, ts._shape_hits = 1000, 500
    assert ts._capture_worth()
