"""ctypes call sites vs the C prototypes of include/kokoro_b200.h, without a GPU: the Python wrappers of the newest entry
points (feature extraction, resampler, decode step, validation metrics) are driven with CPU tensors against a RECORDING
stand-in for libkokoro_b200.so, and every recorded call is checked against the header — argument count, and per
argument the ctypes class the C type needs (a bare Python int for a `long long` parameter would be passed as a 32-bit
int).  Catches wrong order / arity / width before the first hardware run; says nothing about the kernels themselves."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def prototypes():
    header = open(os.path.join(ROOT, "include", "kokoro_b200.h")).read()
    protos = {}
    for m in re.finditer(r"^(?:const\s+)?(?:long long|\w+)\*?\s+(kr_\w+)\s*\(([^;]*?)\);", header, re.M | re.S):
        name, args = m.group(1), re.sub(r"\s+", " ", m.group(2).strip())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = params
    return protos


def expected_ctype(param: str):
    if "*" in param:
        return "pointer"
    t = param.rsplit(" ", 1)[0].replace("const ", "").strip()
    return {"int": ctypes.c_int, "long long": ctypes.c_longlong, "float": ctypes.c_float}[t]


class RecordingLib:
    def __init__(self, returns=None):
        self.calls, self.returns = [], returns or {}

    def __getattr__(self, name):
        if not name.startswith("kr_"):
            raise AttributeError(name)

        class Fn:
            restype = ctypes.c_int

            def __call__(fn, *args):
                self.calls.append((name, args))
                return self.returns.get(name, 0)
        f = Fn()
        object.__setattr__(self, name, f)
        return f


def check_calls(calls):
    protos = prototypes()
    assert calls
    for name, args in calls:
        assert name in protos, f"{name} is not declared in include/kokoro_b200.h"
        params = protos[name]
        assert len(args) == len(params), f"{name}: {len(args)} arguments passed, prototype has {len(params)}: {params}"
        for i, (a, p) in enumerate(zip(args, params)):
            want = expected_ctype(p)
            if want == "pointer":
                ok = (a is None or isinstance(a, ctypes.c_void_p) or type(a).__name__ == "CArgObject"
                      or isinstance(a, ctypes.Array))          # host arrays (pointer tables, block lists)
            else:
                ok = isinstance(a, want)
            assert ok, f"{name}: argument {i} ({p!r}) got {type(a).__name__}"


@pytest.fixture()
def rec(monkeypatch):
    from kokoro_ruslan_b200 import _lib, features, ops
    lib = RecordingLib({"kr_dec_state_size": 256, "kr_val_metrics_acc_floats": 128, "kr_resample_length": 1000,
                        "kr_optim_ctrl_size": 64})
    ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())      # noqa: E731
    for mod in (ops, features):
        monkeypatch.setattr(mod, "lib", lambda: lib)
        monkeypatch.setattr(mod, "_ptr", ptr)
        monkeypatch.setattr(mod, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(ops, "_p", lambda t: None if t is None else t.data_ptr())     # struct-field pointers (kr_gemm_ex)
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(features, "_need_cuda", lambda t, what: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    # nothing is computed in a dry run: hand out zero-filled activations instead of uninitialised memory, so that the few
    # host decisions taken on device results (predicted durations -> expanded length) are deterministic
    from kokoro_ruslan_b200 import engine as engine_mod
    monkeypatch.setattr(engine_mod.AcousticEngine, "_empty",
                        lambda self, *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype))
    return lib


def test_feature_wrappers_match_the_header(rec):
    from kokoro_ruslan_b200 import features
    wav = torch.zeros(2, 6000)
    lens = torch.tensor([6000, 4000])
    features.PitchExtractor.extract_pitch(wav, lengths=lens)
    features.PitchExtractor.extract_pitch(wav[0])
    mel = torch.zeros(2, 30, 80)
    features.EnergyExtractor.extract_energy_from_mel(mel, log_domain=True, frames=torch.tensor([30, 12]))
    features.EnergyExtractor.extract_energy_from_mel(mel.transpose(1, 2).contiguous(), False, channel_major=True, exp_input=True)
    features.resample(wav[:, :1000], 22050, 20506, lengths=torch.tensor([1000, 500]))
    features.speed_perturb(wav[:, :1000], 0.93, lengths=torch.tensor([1000, 500]))
    assert features.trailing_trim_end(mel).shape == (2,) and features.trailing_trim_end(mel[0]).dim() == 0
    tr = features.LogMelSpectrogram(device="cpu")
    tr(wav, lens)
    pipe = features.FeaturePipeline(device="cpu")
    out = pipe(wav, lens)
    assert out["mel_spec"].shape == (2, 80, 24) and out["pitch"].shape == (2, 24) and out["energy"].shape == (2, 24)
    assert out["mel_lengths"].tolist() == [24, 16]
    names = {n for n, _ in rec.calls}
    assert {"kr_pitch_frames", "kr_pitch_track", "kr_energy_frames", "kr_energy_norm", "kr_resample", "kr_resample_length",
            "kr_wave_peak", "kr_mel_stft", "kr_trim_end"} <= names
    check_calls(rec.calls)


def test_decode_backend_and_metrics_match_the_header(rec):
    from kokoro_ruslan_b200 import inference, ops
    D, H, B, Tp, cap = 128, 2, 2, 9, 64

    class Store:
        device = torch.device("cpu")
        pe = torch.zeros(100, D)
        rope_cos = torch.zeros(100, 32)
        rope_sin = torch.zeros(100, 32)
        shadow = torch.zeros(1)

        def p(self, name):
            return torch.zeros(3 * D * D)

        def w(self, name):
            return torch.zeros(D, D, dtype=torch.bfloat16)

        def span(self, buf, first, rows, cols):
            return torch.zeros(rows, cols, dtype=torch.bfloat16)

    eng = type("Eng", (), {"store": Store(), "device": torch.device("cpu"), "D": D})()
    be = inference.CudaDecodeBackend(eng)
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)             # noqa: E731
    bf = torch.bfloat16
    state = z(64, dt=torch.int32)
    be.dec_feed(state, z(B, 80), None, z(D, 80), z(D), be.pe, z(128, D), B, D, 80)
    be.dec_feed(state, z(B, 80), z(B, cap, 80), z(D, 80), z(D), be.pe, z(128, D), B, D, 80)
    qkv = z(128, 3 * D, dt=bf)
    be.dec_attn(state, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], z(64), z(64), z(64), z(B, cap, D, dt=bf),
                z(B, cap, D, dt=bf), -1, None, z(128, D, dt=bf), B, H)
    kv = z(B * Tp, 2 * D, dt=bf).view(B, Tp, 2 * D)
    be.dec_attn(state, z(128, D, dt=bf), None, None, z(64), None, None, kv[:, :, :D], kv[:, :, D:], Tp,
                z(B, Tp, dt=torch.uint8), z(128, D, dt=bf), B, H)
    be.dec_finish(state, z(128, D), z(D), z(D), z(80, D), z(80), z(1, D), z(1), z(B, cap, 80), z(B, 80), z(cap), B, D, 80, cap)
    be.use_gemv, be.state_for_gemv = True, state
    be.gemm(z(128, D, dt=bf), z(3 * D, D, dt=bf), z(128, 3 * D, dt=bf), rows=B)
    be.gemm(z(128, D, dt=bf), z(D, D, dt=bf), z(128, D), bias=z(D), resid=z(128, D), rows=B)
    be.ln_gemv(z(128, D), z(D), z(D), z(4 * D, D, dt=bf), z(4 * D), z(128, 2 * D, dt=bf), B, glu=True)
    gemv = [a for n, a in rec.calls if n == "kr_dec_gemv"]
    assert len(gemv) == 3 and [g[12].value for g in gemv] == [0, 1, 0] and [g[13].value for g in gemv] == [0, 0, 1]
    assert [g[14].value for g in gemv] == [B, B, B] and [g[15].value for g in gemv] == [3 * D, D, 2 * D]
    be.use_gemv = False
    acc = z(ops.val_metrics_acc_floats())
    ops.val_metrics(z(B, 20, 80), z(B, 20, 80), z(B, 20), z(B, 20), torch.tensor([20, 7]), acc)
    ops.val_metrics(z(B, 20, 80), z(B, 20, 80), None, None, torch.tensor([20, 7]), acc)
    ops.average_by_duration(z(B, 20), torch.ones(B, 5, dtype=torch.int64), None, z(B, 20, dt=torch.int32), z(B, 5))
    check_calls(rec.calls)
    # the self-attention call passes the cache strides (row, utterance), the cross call the memory's
    attn = [a for n, a in rec.calls if n == "kr_dec_attn"]
    assert attn[0][13].value == D and attn[0][14].value == cap * D and attn[0][15].value == -1
    assert attn[1][13].value == 2 * D and attn[1][14].value == Tp * 2 * D and attn[1][15].value == Tp


def test_header_parser_sees_every_entry_point():
    protos = prototypes()
    assert len(protos) >= 69 and "kr_gemm_bf16" in protos and "kr_launch_count" in protos
    assert protos["kr_dec_state_size"] == [] and len(protos["kr_pitch_frames"]) == 13


@pytest.mark.parametrize("gemv", [False, True])
def test_inference_engine_dry_run_on_the_recording_lib(rec, monkeypatch, gemv):
    """InferenceEngine.generate end to end on CPU tensors with the recording library: no numerics, but every Python line of
    the device path runs — engine calls in eval mode, geometry tables, buffer shapes / strides asserted by the wrappers,
    cross K/V, DecodeLoop wiring, polling — and every C call it makes is checked against the header."""
    from kokoro_ruslan_b200 import engine as engine_mod
    from kokoro_ruslan_b200 import inference, ops, params
    from kokoro_ruslan_b200.params import ModelConfig
    monkeypatch.setenv("KR_DECODE_GRAPH", "0")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    for mod in (engine_mod, params):
        if hasattr(mod, "lib"):
            monkeypatch.setattr(mod, "lib", lambda: rec)
    frames = {"n": 0}

    class Finish:
        restype = ctypes.c_int

        def __call__(self, *args):
            rec.calls.append(("kr_dec_finish", args))
            st = (ctypes.c_int * 8).from_address(args[0].value)      # the stand-in "device" advances the state
            if st[1]:
                return 0                                             # done: the real kernel returns early too
            st[0] += 1
            frames["n"] = st[0]
            if st[0] >= st[3] + 2:                                   # two frames past the minimum length
                st[1], st[2] = 1, st[0]
            return 0
    object.__setattr__(rec, "kr_dec_finish", Finish())
    cfg = ModelConfig(vocab_size=59, mel_dim=80, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                      n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64,
                      n_variance_bins=256)
    eng = engine_mod.AcousticEngine(cfg, device="cpu", with_ema=False, multi_stream=False)
    inf = inference.InferenceEngine(eng)
    assert inf.be.use_gemv                  # the product default (<= 8 utterances); False = the path 9..16 utterances take
    inf.be.use_gemv = gemv
    idx = torch.randint(1, 59, (2, 11))
    dur = torch.randint(1, 4, (2, 11))
    dur[1, 8:] = 0
    mel, probs = inf.generate(idx, torch.randint(0, 3, (2, 11)), durations=dur, return_stop_probs=True)
    Tp = int(dur.sum(dim=1).max())
    lo, hi = inference.generation_bounds(Tp, 400)
    assert mel.shape == (2, lo + 2, 80) and probs.shape == (lo + 2,)
    assert frames["n"] == lo + 2                                      # polling stopped calling once `done` was set ...
    n_finish = sum(1 for n, _ in rec.calls if n == "kr_dec_finish")
    assert n_finish == -(-(lo + 2) // 32) * 32                       # ... at the next multiple of the poll interval
    names = [n for n, _ in rec.calls]
    per_step = 2 + 2 * cfg.n_decoder_layers                          # feed + finish + 2 attentions per layer
    assert names.count("kr_dec_attn") == 2 * cfg.n_decoder_layers * n_finish
    assert names.count("kr_dec_feed") == n_finish and per_step == 6
    # kr_dec_gemv path: six skinny projections per layer and no separate LayerNorm / GLU launches inside the loop
    assert names.count("kr_dec_gemv") == (6 * cfg.n_decoder_layers * n_finish if gemv else 0)
    assert (names.count("kr_glu_fwd") == cfg.n_encoder_layers) == gemv          # only the encoder FFNs are left
    # eval mode: no dropout specs reached the kernels, the engine's training flag is restored
    assert eng.training is True
    check_calls([c for c in rec.calls if c[0] in ("kr_dec_feed", "kr_dec_attn", "kr_dec_finish", "kr_dec_gemv", "kr_lr_index",
                                                   "kr_expand_adapt", "kr_embed_fwd", "kr_layernorm_fwd", "kr_glu_fwd",
                                                   "kr_rmsnorm_resid_fwd", "kr_eq_mask_i64", "kr_scatter_rows",
                                                   "kr_vp_head_fwd", "kr_gn_fwd", "kr_memset_zero", "kr_gemm_bf16",
                                                   "kr_qkv_prep_fwd", "kr_attn_fwd")])
    # predicted-duration path (no override): zeros from the stand-in -> Tp is padded to the 3-frame minimum
    mem, fmask, log_dur, Tp0 = inf.encode_and_expand(idx, None)
    assert Tp0 == 3 and mem.shape == (2 * 3, 128) and fmask.shape == (2, 3) and log_dur.shape == (2, 11)


def test_eval_losses_with_metrics_dry_run(rec, monkeypatch):
    """TrainStep.eval_losses(metrics=acc) on the recording library: the EMA-weights eval forward, the loss call and the
    metrics fold run through every Python line of the device path; the C calls are checked against the header."""
    monkeypatch.setenv("KR_STREAMS", "0")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    from kokoro_ruslan_b200 import engine as engine_mod
    from kokoro_ruslan_b200 import optim, params
    from kokoro_ruslan_b200 import train_step as ts_mod
    from kokoro_ruslan_b200.params import ModelConfig
    from oracle import acoustic as oa
    ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())      # noqa: E731
    for mod in (engine_mod, params, optim, ts_mod):
        for name, val in (("lib", lambda: rec), ("_ptr", ptr), ("_stream", lambda: ctypes.c_void_p(0))):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, val)
    cfg = ModelConfig(vocab_size=59, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256, n_decoder_layers=2,
                      decoder_ff_dim=256, max_decoder_seq_len=1200, variance_filter_size=64)
    ts = ts_mod.TrainStep(cfg, sched_cfg=ts_mod.ScheduleConfig(total_steps=10), device="cpu", use_graphs=False)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    acc = ts.new_val_metrics()
    assert acc.shape == (128,) and acc.dtype == torch.float32
    before = (ts.engine.store.params.data_ptr(), ts.engine.training)
    losses = ts.eval_losses(batch, use_ema=True, metrics=acc)
    assert losses.shape == (6,)
    assert (ts.engine.store.params.data_ptr(), ts.engine.training) == before      # EMA swap and eval mode undone
    assert set(ts.read_val_metrics(acc)) == {"val_spectral_convergence", "val_f0_rmse"}
    names = [n for n, _ in rec.calls]
    assert names.count("kr_val_metrics") == 1 and names.count("kr_losses_fwd_bwd") == 1
    vm = [a for n, a in rec.calls if n == "kr_val_metrics"][0]
    assert (vm[6].value, vm[7].value, vm[9].value) == (3, 150, 80)                # B, T, n_mels
    check_calls([c for c in rec.calls if c[0] in ("kr_val_metrics", "kr_losses_fwd_bwd", "kr_cast_bf16", "kr_expand_adapt")])


@pytest.mark.parametrize("with_dropout", [False, True])
def test_full_training_step_dry_run_conforms_to_the_header(rec, monkeypatch, with_dropout):
    """One whole optimizer step (forward, losses, backward, norms, step control, AdamW / EMA, projection) on the recording
    library: a regression net over every ctypes call site of the training path — ~40 entry points, a few hundred calls —
    against the header prototypes."""
    monkeypatch.setenv("KR_STREAMS", "0")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    from kokoro_ruslan_b200 import engine as engine_mod
    from kokoro_ruslan_b200 import optim, params
    from kokoro_ruslan_b200 import train_step as ts_mod
    from kokoro_ruslan_b200.engine import DropoutConfig
    from kokoro_ruslan_b200.params import ModelConfig
    from oracle import acoustic as oa
    ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())      # noqa: E731
    for mod in (engine_mod, params, optim, ts_mod):
        for name, val in (("lib", lambda: rec), ("_ptr", ptr), ("_stream", lambda: ctypes.c_void_p(0))):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, val)
    cfg = ModelConfig(vocab_size=59, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256, n_decoder_layers=2,
                      decoder_ff_dim=256, max_decoder_seq_len=1200, variance_filter_size=64)
    ts = ts_mod.TrainStep(cfg, sched_cfg=ts_mod.ScheduleConfig(total_steps=10), device="cpu", use_graphs=False,
                          dropout=DropoutConfig.reference_training() if with_dropout else None)
    rec.calls.clear()
    losses = ts.train_step(oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True))
    assert losses.shape == (6,)
    names = {n for n, _ in rec.calls}
    assert {"kr_gemm_bf16", "kr_attn_fwd", "kr_attn_bwd", "kr_losses_fwd_bwd", "kr_adamw_step", "kr_step_control",
            "kr_expand_adapt", "kr_lr_index", "kr_layernorm_bwd", "kr_qkv_prep_bwd"} <= names
    assert ("kr_drop_begin" in names) == with_dropout
    assert len(rec.calls) > 200 and len(names) >= 35
    check_calls(rec.calls)


def test_model_forward_inference_and_synthesizer_dry_run(rec, monkeypatch):
    """KokoroModel.forward_inference (reference signature) -> HiFi-GAN through Synthesizer on the recording library: the
    Python of the whole text-free TTS path, ~2000 C calls checked against the header (incl. kr_gemm_ex's argument struct)."""
    monkeypatch.setenv("KR_STREAMS", "0")
    monkeypatch.setenv("KR_DECODE_GRAPH", "0")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: None)
    from kokoro_ruslan_b200 import engine as engine_mod
    from kokoro_ruslan_b200 import hifigan, inference, model, optim, params
    ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())      # noqa: E731
    for mod in (engine_mod, params, optim, hifigan, model, inference):
        for name, val in (("lib", lambda: rec), ("_ptr", ptr), ("_stream", lambda: ctypes.c_void_p(0))):
            if hasattr(mod, name):
                monkeypatch.setattr(mod, name, val)

    class Finish:
        restype = ctypes.c_int

        def __call__(self, *args):
            rec.calls.append(("kr_dec_finish", args))
            st = (ctypes.c_int * 8).from_address(args[0].value)
            if not st[1]:
                st[0] += 1
                if st[0] >= st[3] + 2:
                    st[1], st[2] = 1, st[0]
            return 0
    object.__setattr__(rec, "kr_dec_finish", Finish())
    with pytest.raises(NotImplementedError):
        model.KokoroModel(vocab_size=59, device="cpu")              # the reference's qk_norm=False default is not on this path
    m = model.KokoroModel(vocab_size=59, qk_norm=True, hidden_dim=128, n_encoder_layers=1, n_heads=2, encoder_ff_dim=256,
                          n_decoder_layers=1, decoder_ff_dim=256, max_decoder_seq_len=400, variance_filter_size=64,
                          device="cpu")
    m.eval()
    idx = torch.randint(1, 59, (1, 12))
    mel = m.forward_inference(idx, min_len_floor=48, max_len_cap=64, stress_indices=torch.zeros(1, 12, dtype=torch.int64))
    assert mel.shape == (1, 50, 80)
    # forward(mel_specs=None) dispatches to forward_inference like the reference (model.py:813-818)
    assert m(idx).shape == (1, 14, 80)                               # lo = 12 -> the stand-in stops two frames later
    # nn.Module surface the reference trainer uses (trainer.py:835, 845-881)
    import copy
    info = m.get_model_info()
    assert info["n_decoder_layers"] == 1 and info["total_parameters"] == sum(p.numel() for p in m.parameters())
    assert m.variance_adaptor.pitch_bins.shape == (255,)
    mods = dict(m.named_modules())
    assert mods["decoder.layers.0.ff.linear1"].weight.shape == (512, 128)
    assert mods["transformer_encoder_layers.0.ff.linear2"].weight is dict(m.named_parameters())[
        "transformer_encoder_layers.0.ff.linear2.weight"]
    assert [n for n, _ in m.named_parameters()] == m._names and len(list(m.modules())) > 50
    assert list(m.state_dict().keys())[:3] == ["text_embedding.weight", "stress_embedding.weight", "positional_encoding.pe"]
    twin = copy.deepcopy(m)
    assert twin is not m and twin.engine is not m.engine and not twin.training
    assert all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(), m.state_dict().values()))
    twin.state_dict()["text_embedding.weight"].mul_(0.0)            # in-place EMA-style update reaches the flat buffer ...
    assert float(twin.engine.store.p("text_embedding.weight").abs().max()) == 0.0
    assert float(m.engine.store.p("text_embedding.weight").abs().max()) > 0.0      # ... of the copy only
    voc = hifigan.HiFiGANGenerator(hifigan.HiFiGANConfig.get_default_config(), device="cpu", use_graphs=False)
    audio, mel = inference.Synthesizer(m, voc)(idx, min_len_floor=48, max_len_cap=64)
    assert audio.shape == (1, 50 * 256) and mel.shape == (1, 50, 80)
    assert len(rec.calls) > 1000
    check_calls(rec.calls)
