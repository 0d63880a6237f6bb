"""Autoregressive inference on the device: InferenceEngine / DecodeLoop with CudaDecodeBackend
(kr_dec_feed / kr_dec_attn / kr_dec_finish + the training-path kernels on a 128-row padded batch, one CUDA graph per
step) against the oracle's forward_inference, which is pinned to the live reference.

Conditioning of the comparison: at inference the pitch / energy embeddings are selected by bucketising PREDICTED values
into 256 bins of width 1/255, which is below the bf16 noise of the predictors (the reference's own autocast run has the
same property), so a neighbouring bin is routinely selected.  With i.i.d. random embedding tables a neighbouring row is
an unrelated vector and the comparison would measure nothing but that (first hardware run of round 2: 47 % of the
memory rows "differ"; the fp32 oracle with 3e-3 relative noise on its own predictions gives 52 %).  The fixtures
therefore use embedding tables that vary smoothly with the bin index — as trained tables do — and the test checks the
three stages separately: predicted values within tolerance, selected bins within the bins that tolerance spans, memory
rows within tolerance."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
HERE = os.path.dirname(os.path.abspath(__file__))


def _setup():
    from kokoro_ruslan_b200.engine import AcousticEngine
    from kokoro_ruslan_b200.inference import InferenceEngine
    from kokoro_ruslan_b200.params import ModelConfig
    from oracle import acoustic as oa
    f = np.load(os.path.join(HERE, "golden", "inference.npz"))
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    sd = oa.seeded_state_dict(ocfg, seed=int(f["seed"]))
    sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([float(f["dur_bias"])])
    sd["stop_token_predictor.bias"] = torch.tensor([float(f["stop_bias"])])
    _smooth_variance_embeddings(sd)
    cfg = ModelConfig(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                      n_encoder_layers=ocfg.n_encoder_layers, n_heads=ocfg.n_heads, encoder_ff_dim=ocfg.ff_dim,
                      n_decoder_layers=ocfg.n_decoder_layers, decoder_ff_dim=ocfg.ff_dim,
                      max_decoder_seq_len=ocfg.max_len, variance_filter_size=ocfg.variance_filter,
                      n_variance_bins=ocfg.n_bins)
    eng = AcousticEngine(cfg, device="cuda:0", with_ema=False)
    eng.store.load_state_dict(sd)
    return f, ocfg, sd, InferenceEngine(eng)


def _smooth_variance_embeddings(sd, sigma_bins: float = 24.0):
    """Low-pass the pitch / energy embedding tables along the bin axis (same RMS): neighbouring bins -> neighbouring
    vectors, see the module docstring."""
    va = "duration_adaptor.variance_adaptor."
    for k in (va + "pitch_embedding.weight", va + "energy_embedding.weight"):
        w = sd[k]
        n = w.shape[0]
        i = torch.arange(n, dtype=torch.float32)
        kern = torch.exp(-0.5 * ((i[:, None] - i[None, :]) / sigma_bins) ** 2)
        kern = kern / kern.sum(dim=1, keepdim=True)
        sm = kern @ w
        sd[k] = (sm * (w.pow(2).mean().sqrt() / sm.pow(2).mean().sqrt())).contiguous()


def _oracle_durations(sd, ocfg, idx, stress):
    from oracle import inference as oi
    _, _, log_dur = oi.encode_and_expand(sd, ocfg, idx, stress)
    return torch.clamp(torch.round(torch.expm1(log_dur)), min=0).long()


def test_encode_and_expand_matches_oracle():
    from oracle import inference as oi
    f, ocfg, sd, inf = _setup()
    idx, stress = torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"])
    det = {}
    want_mem, want_pad, want_ld = oi.encode_and_expand(sd, ocfg, idx, stress, details=det)
    dur = _oracle_durations(sd, ocfg, idx, stress)
    mem, fmask, log_dur, Tp = inf.encode_and_expand(idx.cuda(), stress.cuda(), durations=dur)
    assert Tp == want_mem.shape[1]
    assert torch.equal(fmask.cpu().bool(), want_pad)
    assert float((log_dur.cpu() - want_ld).abs().max()) < 2e-2 * max(1.0, float(want_ld.abs().max()))
    live = ~want_pad
    got = inf.last_predictions
    enc_err = float((got["encoder"].float().cpu().view_as(det["encoder"]) - det["encoder"]).abs().max()) \
        / float(det["encoder"].abs().max())
    assert enc_err < 1e-2, enc_err
    for name in ("pitch", "energy"):
        w, g = det[name], got[name].float().cpu().view_as(det[name])
        tol = 2e-2 * max(1.0, float(w[live].abs().max()))
        err = float((g - w)[live].abs().max())
        assert err < tol, (name, err, tol)
        # selected bins: within the number of 1/255-wide bins the prediction tolerance spans (values clamp to [0, 1])
        wb, gb = det[name + "_idx"], got[name + "_bins"].cpu().long().view_as(det[name + "_idx"])
        dbin = int((gb - wb)[live].abs().max())
        assert dbin <= int(err * 255) + 1, (name, dbin, err)
        print(f"{name}: max err {err:.2e}, max bin distance {dbin}")
    got_mem = mem.float().cpu().view(1, Tp, -1)
    err = float((got_mem - want_mem).abs().max()) / float(want_mem.abs().max())
    print(f"encoder {enc_err:.2e}, expanded memory {err:.2e}")
    assert err < 2e-2, err
    # predicted durations on their own: within one frame per token of the oracle's
    _, _, _, Tp_free = inf.encode_and_expand(idx.cuda(), stress.cuda())
    assert abs(Tp_free - Tp) <= idx.shape[1]


def test_teacher_forced_decode_matches_oracle():
    from oracle import inference as oi
    f, ocfg, sd, inf = _setup()
    idx, stress = torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"])
    want, want_p, raw = oi.forward_inference(sd, ocfg, idx, stress, return_raw=True)
    n = want.shape[1]
    forced = torch.zeros(1, 1600, ocfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]
    got, probs = inf.generate(idx.cuda(), stress.cuda(), forced=forced.cuda(), return_stop_probs=True,
                              durations=_oracle_durations(sd, ocfg, idx, stress))
    assert abs(got.shape[1] - n) <= 1, (got.shape, want.shape)
    m = min(got.shape[1], n)
    err = float((got[:, :m].cpu() - want[:, :m]).abs().max()) / float(want.abs().max())
    assert err < 2e-2, err
    assert float((probs[:m].cpu() - torch.tensor(want_p)[:m]).abs().max()) < 3e-2


def test_free_running_batch_generation():
    from oracle import inference as oi
    f, ocfg, sd, inf = _setup()
    idx = torch.from_numpy(f["idx2"])
    want = oi.forward_inference(sd, ocfg, idx, None, stop_threshold=0.45)
    got = inf.generate(idx.cuda(), None, stop_threshold=0.45, durations=_oracle_durations(sd, ocfg, idx, None)).cpu()
    assert got.shape[0] == 2 and abs(got.shape[1] - want.shape[1]) <= 3, (got.shape, want.shape)
    assert bool(torch.isfinite(got).all()) and float(got.max()) <= 2.0 and float(got.min()) >= -11.5
    assert float((got[:, :3] - want[:, :3]).abs().max()) / float(want.abs().max()) < 4e-2


def test_model_forward_inference_surface_and_graph_equals_eager(monkeypatch):
    """KokoroModel.forward_inference (reference signature) end to end at the full model width, and the CUDA-graph replay
    against the same loop launched eagerly (KR_DECODE_GRAPH=0): identical frames."""
    from kokoro_ruslan_b200.model import KokoroModel
    m = KokoroModel(vocab_size=59, encoder_ff_dim=1536, decoder_ff_dim=1536, qk_norm=True)
    m.eval()
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(1, 59, (1, 16), generator=g).cuda()
    a = m.forward_inference(idx, max_len_cap=96)
    assert a.dim() == 3 and a.shape[0] == 1 and a.shape[2] == 80 and 12 <= a.shape[1] <= 96
    monkeypatch.setenv("KR_DECODE_GRAPH", "0")
    b = m.forward_inference(idx, max_len_cap=96)
    assert a.shape == b.shape and torch.equal(a, b)


def test_synthesizer_mel_to_waveform():
    """forward_inference -> HiFi-GAN on the device: 256 samples per generated frame, finite audio in [-1, 1]."""
    from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
    from kokoro_ruslan_b200.inference import Synthesizer
    from kokoro_ruslan_b200.model import KokoroModel
    m = KokoroModel(vocab_size=59, encoder_ff_dim=1536, decoder_ff_dim=1536, qk_norm=True)
    m.eval()
    voc = HiFiGANGenerator(HiFiGANConfig.get_default_config())
    idx = torch.randint(1, 59, (1, 12), generator=torch.Generator().manual_seed(5)).cuda()
    audio, mel = Synthesizer(m, voc)(idx, min_len_floor=48, max_len_cap=64)      # 49..64 frames
    assert audio.shape == (1, mel.shape[1] * 256) and bool(torch.isfinite(audio).all())
    assert float(audio.abs().max()) <= 1.0


def test_padded_gemm_projection_path_equals_gemv_path():
    """The two projection paths of the decode step — kr_dec_gemv (<= 8 utterances, the default) and the tcgen05 GEMM on the
    128-row padded buffers (9..16 utterances) — give the same teacher-forced frames, and both match the oracle."""
    from oracle import inference as oi
    f, ocfg, sd, inf = _setup()
    idx = torch.from_numpy(f["idx2"])
    want, _, raw = oi.forward_inference(sd, ocfg, idx, None, stop_threshold=0.45, return_raw=True)
    n = want.shape[1]
    forced = torch.zeros(2, 1600, ocfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]
    dur = _oracle_durations(sd, ocfg, idx, None)
    assert inf.be.use_gemv
    b = inf.generate(idx.cuda(), None, stop_threshold=0.45, forced=forced.cuda(), durations=dur).cpu()
    from kokoro_ruslan_b200.inference import InferenceEngine
    inf2 = InferenceEngine(inf.eng)
    inf2.be.use_gemv = False
    a = inf2.generate(idx.cuda(), None, stop_threshold=0.45, forced=forced.cuda(), durations=dur).cpu()
    assert a.shape == b.shape
    assert float((a - b).abs().max()) / float(want.abs().max()) < 1.2e-2      # GLU input not rounded to bf16 in the fused path
    m = min(a.shape[1], n)
    for got in (a, b):
        assert float((got[:, :m] - want[:, :m]).abs().max()) / float(want.abs().max()) < 2e-2
