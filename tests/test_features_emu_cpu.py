"""Host emulation of the pitch / energy feature kernels (kokoro_ruslan_b200/csrc/kr_features_core.cuh compiled by g++
with -DKR_HOST_EMU, one "thread" per block; tests/emu/features_emu.cpp loops over the grid like the launch wrappers)
against the LIVE-reference fixtures of tests/golden/features.npz and the numpy oracle.  This checks the kernels' own
source — index arithmetic, in-place FFT pair, CMND, every thresholded decision, rank-counting quantiles — on the CPU;
the -m gpu test (tests/test_features_gpu.py) checks the same entry points on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=["seq", "simt"])
def emu(request, tmp_path_factory):
    """The kernels' bodies as a host library: "seq" = one sequential thread per block; "simt" = one host thread per CUDA
    thread with the real block / warp geometry, barriers and warp reductions (tests/emu/emu_simt.h)."""
    simt = request.param == "simt"
    so = tmp_path_factory.mktemp("emu") / ("features_emu_%s.so" % request.param)
    flags = ["-DKR_HOST_EMU_SIMT", "-pthread", "-I", os.path.join(HERE, "emu")] if simt else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-x", "c++", "-I",
                    os.path.join(ROOT, "kokoro_ruslan_b200", "csrc"), os.path.join(HERE, "emu", "features_emu.cpp"), "-o", str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    lib.simt = simt
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def emu_pitch(lib, wav, lengths=None, sr=22050, fmin=50.0, fmax=800.0):
    wav = np.ascontiguousarray(np.atleast_2d(wav), dtype=np.float32)
    B, N = wav.shape
    lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int64)
    T = max(lib.emu_pitch_num_frames(ctypes.c_longlong(int(n))) for n in ([N] if lens is None else lens))
    bufs = [np.full((B, T), np.nan, np.float32) for _ in range(3)]
    work, out = np.zeros((B, T), np.float32), np.full((B, T), np.nan, np.float32)
    f = ctypes.c_float
    assert lib.emu_pitch_frames(_p(wav), _p(lens), *map(_p, bufs), B, ctypes.c_longlong(N), T, sr, f(fmin), f(fmax)) == 0
    assert lib.emu_pitch_track(*map(_p, bufs), _p(lens), _p(work), _p(out), B, ctypes.c_longlong(N), T, f(fmin), f(fmax)) == 0
    return out


def emu_energy(lib, mel, time_major=True, exp_input=False, log_domain=True, frames=None):
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    B, T, M = mel.shape if time_major else (mel.shape[0], mel.shape[2], mel.shape[1])
    e, out = np.zeros((B, T), np.float32), np.full((B, T), np.nan, np.float32)
    fr = None if frames is None else np.ascontiguousarray(frames, dtype=np.int64)
    assert lib.emu_energy_frames(_p(mel), _p(e), B, T, M, int(time_major), int(exp_input), int(log_domain)) == 0
    assert lib.emu_energy_norm(_p(e), _p(fr), _p(out), B, T) == 0
    return out


def test_pitch_kernels_match_live_reference(emu):
    f = np.load(os.path.join(HERE, "golden", "features.npz"))
    got = emu_pitch(emu, f["wav"])
    assert got.shape == f["pitch"].shape
    d = np.abs(got - f["pitch"])
    assert (d < 1e-4).mean() >= 0.99 and d.max() < 0.05, ((d < 1e-4).mean(), d.max())
    assert ((got > 0) == (f["pitch"] > 0)).mean() >= 0.99
    short = emu_pitch(emu, f["wav"][0, :1500])[0]                      # shorter than one analysis window
    assert short.shape == f["pitch_short"].shape and np.abs(short - f["pitch_short"]).max() < 1e-4


def test_pitch_ragged_batch_equals_per_item(emu):
    """Per-utterance lengths: every row equals the utterance processed alone, padding frames are zero."""
    from oracle import features as of
    f = np.load(os.path.join(HERE, "golden", "features.npz"))
    lens = np.array([35280, 20000, 9001, 1500])
    wav = f["wav"].copy()
    for b, n in enumerate(lens):
        wav[b, n:] = 7.0                                               # garbage beyond the length must never be read
    got = emu_pitch(emu, wav, lens)
    for b, n in enumerate(lens):
        single = emu_pitch(emu, f["wav"][b, :n])[0]
        T = single.shape[0]
        assert np.array_equal(got[b, :T], single)
        assert np.all(got[b, T:] == 0.0)
        want = of.extract_pitch(f["wav"][b, :n])
        d = np.abs(single - want)
        assert (d < 1e-4).mean() >= 0.98 and d.max() < 0.05, (b, (d < 1e-4).mean(), d.max())


def test_energy_kernels_match_live_reference(emu):
    f = np.load(os.path.join(HERE, "golden", "features.npz"))
    mel = f["mel"]
    assert np.abs(emu_energy(emu, mel, log_domain=True) - f["e_log"]).max() < 1e-5
    assert np.abs(emu_energy(emu, np.exp(mel), log_domain=False) - f["e_lin"]).max() < 1e-5
    # what the product pipeline does: log-mel in (the mel-STFT kernel's output layout), linear-power semantics out
    cm = np.ascontiguousarray(mel.transpose(0, 2, 1))
    assert np.abs(emu_energy(emu, cm, time_major=False, exp_input=True, log_domain=False) - f["e_lin"]).max() < 1e-5
    assert np.abs(emu_energy(emu, mel[:, :2], log_domain=True) - f["e_short"]).max() < 1e-6    # < 3 frames: min / max
    # ragged: normalisation statistics over the valid frames only, zeros beyond
    from oracle import features as of
    got = emu_energy(emu, mel, log_domain=True, frames=[140, 77, 2])
    for b, t in enumerate([140, 77, 2]):
        assert np.abs(got[b, :t] - of.extract_energy_from_mel(mel[b, :t], True)).max() < 1e-5
        assert np.all(got[b, t:] == 0.0)


def test_rank_counting_quantiles_with_ties(emu):
    """Order statistics by rank counting must behave like a sort when values repeat (silences give many equal frames)."""
    from oracle import features as of
    rng = np.random.default_rng(0)
    e = rng.integers(0, 5, size=(6, 1, 57)).astype(np.float32)        # heavy ties
    mel = np.repeat(e.transpose(0, 2, 1), 4, axis=2)                   # (6, 57, 4): the frame mean is e itself
    got = emu_energy(emu, mel, log_domain=True)
    want = np.stack([of.extract_energy_from_mel(mel[b], True) for b in range(6)])
    assert np.abs(got - want).max() < 1e-6


def test_radix4_mel_stft_variant_matches_torchaudio_golden(emu):
    """kr_mel_stft kernel body (radix-4 FFT + digit reversal) against torchaudio's own output and the float64 oracle,
    at the gates of the device test of the default kernel (tests/test_melstft_gpu.py): 1e-3 absolute on the log-mel."""
    import torch
    from kokoro_ruslan_b200.features import mel_filterbank_htk
    from oracle import melstft as om
    fix = np.load(os.path.join(HERE, "golden", "melstft.npz"))
    fb_t = np.ascontiguousarray(mel_filterbank_htk(513, 0.0, 8000.0, 80, 22050).t().numpy())
    # [first non-zero bin, one past the last) per filter, as features.LogMelSpectrogram builds it
    nz = fb_t != 0
    rng = np.ascontiguousarray(np.stack([nz.argmax(axis=1), 513 - nz[:, ::-1].argmax(axis=1)], axis=1).astype(np.int32))
    assert int((rng[:, 1] - rng[:, 0]).max()) < 64 and bool(nz.any(axis=1).all())
    for case in ("a", "b", "c"):
        wav = np.ascontiguousarray(fix[f"wav_{case}"], dtype=np.float32).reshape(1, -1)
        want = fix[f"mel_{case}"]
        n = wav.shape[1]
        frames = 1 + n // 256
        peak = np.array([np.abs(wav).max()], np.float32)
        out = np.full((1, 80, frames), np.nan, np.float32)
        assert emu.emu_mel_stft_r4(_p(wav), None, _p(peak), _p(fb_t), _p(rng), _p(out), 1, ctypes.c_longlong(n), frames, 80,
                                   ctypes.c_float(1e-9)) == 0
        dense = np.full((1, 80, frames), np.nan, np.float32)        # skipping exact zeros changes no bit
        assert emu.emu_mel_stft_r4(_p(wav), None, _p(peak), _p(fb_t), None, _p(dense), 1, ctypes.c_longlong(n), frames, 80,
                                   ctypes.c_float(1e-9)) == 0
        assert np.array_equal(out, dense)
        assert out[0].shape == want.shape
        assert np.abs(out[0] - want).max() < 1e-3, case
        assert np.abs(out[0] - om.log_mel(wav[0])).max() < 1e-3
    # ragged: frames beyond 1 + len // 256 are zero, the others equal the single run
    wav2 = np.zeros((2, 9000), np.float32)
    wav2[0] = fix["wav_a"][:9000]
    wav2[1, :5000] = fix["wav_a"][:5000]
    lens = np.array([9000, 5000], np.int64)
    out2 = np.full((2, 80, 36), np.nan, np.float32)
    assert emu.emu_mel_stft_r4(_p(wav2), _p(lens), None, _p(fb_t), _p(rng), _p(out2), 2, ctypes.c_longlong(9000), 36, 80,
                               ctypes.c_float(1e-9)) == 0
    single = np.full((1, 80, 20), np.nan, np.float32)
    w1 = np.ascontiguousarray(wav2[1:2, :5000])
    assert emu.emu_mel_stft_r4(_p(w1), None, None, _p(fb_t), _p(rng), _p(single), 1, ctypes.c_longlong(5000), 20, 80,
                               ctypes.c_float(1e-9)) == 0
    assert np.array_equal(out2[1, :, :20], single[0]) and np.all(out2[1, :, 20:] == 0.0)


def _reference_trim_end(mel):
    """inference/inference.py:593-619, literally (mel: (T, n_mels) torch tensor)."""
    import torch
    frame_means = mel.mean(dim=-1)
    q10 = float(torch.quantile(frame_means, 0.10).item())
    q20 = float(torch.quantile(frame_means, 0.20).item())
    thr = max(-9.8, min(-9.2, 0.5 * (q10 + q20)))
    voiced = (frame_means > thr).nonzero(as_tuple=False).squeeze(-1)
    if voiced.numel() == 0:
        return mel.shape[0]
    proposed_end = min(mel.shape[0], int(voiced[-1]) + 24 + 1)
    return min(max(60, proposed_end), mel.shape[0])


def test_trailing_trim_matches_reference_rule(emu):
    import torch
    g = torch.Generator().manual_seed(0)
    cases = []
    for T, speech_end in ((300, 180), (300, 299), (90, 20), (40, 10), (500, 0), (200, 100)):
        mel = torch.full((T, 80), -10.5) + 0.3 * torch.randn(T, 80, generator=g)          # quiet tail level
        mel[:speech_end] = -5.0 + 2.0 * torch.randn(speech_end, 80, generator=g)          # speech
        cases.append(mel.clamp(-11.5, 2.0))
    cases.append(torch.full((120, 80), -11.0))                                            # nothing above the threshold
    for mel in cases:
        T = mel.shape[0]
        m = np.ascontiguousarray(mel.numpy()[None])
        e, t_end = np.zeros((1, T), np.float32), np.full(1, -1, np.int32)
        assert emu.emu_energy_frames(_p(m), _p(e), 1, T, 80, 1, 0, 1) == 0
        assert emu.emu_trim_end(_p(e), None, _p(t_end), 1, T) == 0
        assert int(t_end[0]) == _reference_trim_end(mel), (T, int(t_end[0]), _reference_trim_end(mel))
    assert _reference_trim_end(cases[0]) == 180 + 24 and _reference_trim_end(cases[3]) == 40


def test_edge_inputs_match_live_reference(emu):
    """Silence, DC, one sample, noise, an impulse, tones at both ends of the pitch range; energy on 1-5 frames."""
    from oracle import features as of
    f = np.load(os.path.join(HERE, "golden", "features_edge.npz"))
    for k in ("zeros", "dc", "n1", "noise", "impulse", "tone60", "tone790"):
        got, want = emu_pitch(emu, f[f"wav_{k}"])[0], f[f"pitch_{k}"]
        assert got.shape == want.shape and not np.isnan(got).any(), k
        assert np.abs(got - want).max() < 1e-5, (k, np.abs(got - want).max())
        assert np.abs(of.extract_pitch(f[f"wav_{k}"]) - want).max() < 1e-5, k           # the oracle on the same inputs
    assert (f["pitch_tone60"] > 0).all() and (f["pitch_noise"] == 0).all()
    for T in (1, 2, 3, 5):
        got = emu_energy(emu, f[f"mel_{T}"][None], log_domain=True)[0]
        assert np.abs(got - f[f"energy_{T}"]).max() < 1e-6, T
