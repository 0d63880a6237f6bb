"""Pitch / energy feature kernels on the device against the LIVE-reference fixtures and the numpy
oracle, through the C ABI (kr_pitch_frames / kr_pitch_track / kr_energy_frames / kr_energy_norm).

The same kernel source is also verified on the CPU by the host emulation (tests/test_features_emu_cpu.py: identical to
the live reference on every fixture frame); first hardware run: round 2, all green."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
HERE = os.path.dirname(os.path.abspath(__file__))


def _fix():
    return np.load(os.path.join(HERE, "golden", "features.npz"))


def test_pitch_matches_live_reference():
    from kokoro_ruslan_b200.features import PitchExtractor
    f = _fix()
    got = PitchExtractor.extract_pitch(torch.from_numpy(f["wav"]).cuda()).cpu().numpy()
    assert got.shape == f["pitch"].shape
    d = np.abs(got - f["pitch"])
    assert (d < 1e-4).mean() >= 0.99 and d.max() < 0.05, ((d < 1e-4).mean(), d.max())
    assert ((got > 0) == (f["pitch"] > 0)).mean() >= 0.99
    short = PitchExtractor.extract_pitch(torch.from_numpy(f["wav"][0, :1500]).cuda()).cpu().numpy()
    assert short.shape == f["pitch_short"].shape and np.abs(short - f["pitch_short"]).max() < 1e-4


def test_pitch_ragged_batch_equals_per_item():
    from kokoro_ruslan_b200.features import PitchExtractor
    f = _fix()
    lens = [35280, 20000, 9001, 1500]
    wav = torch.from_numpy(f["wav"].copy())
    for b, n in enumerate(lens):
        wav[b, n:] = 7.0
    got = PitchExtractor.extract_pitch(wav.cuda(), lengths=torch.tensor(lens)).cpu()
    for b, n in enumerate(lens):
        single = PitchExtractor.extract_pitch(torch.from_numpy(f["wav"][b, :n]).cuda()).cpu()
        T = single.shape[0]
        assert torch.equal(got[b, :T], single) and float(got[b, T:].abs().sum()) == 0.0


def test_energy_matches_live_reference():
    from kokoro_ruslan_b200.features import EnergyExtractor
    f = _fix()
    mel = torch.from_numpy(f["mel"]).cuda()
    ex = EnergyExtractor.extract_energy_from_mel
    assert float((ex(mel).cpu() - torch.from_numpy(f["e_log"])).abs().max()) < 1e-5              # heuristic -> log branch
    assert float((ex(mel.exp(), log_domain=False).cpu() - torch.from_numpy(f["e_lin"])).abs().max()) < 1e-5
    cm = mel.transpose(1, 2).contiguous()
    got = ex(cm, log_domain=False, channel_major=True, exp_input=True).cpu()
    assert float((got - torch.from_numpy(f["e_lin"])).abs().max()) < 1e-5
    assert float((ex(mel[:, :2]).cpu() - torch.from_numpy(f["e_short"])).abs().max()) < 1e-6
    assert float((ex(mel[0]).cpu() - torch.from_numpy(f["e_log"][0])).abs().max()) < 1e-5         # (frames, n_mels) input


def test_feature_pipeline_full_size():
    """BASELINE shape (8 utterances x 800 frames): the batched pipeline equals the per-item oracle composition
    (dataset.py:672-815: peak normalisation -> log-mel, pitch of the normalised audio, energy of the linear mel)."""
    from kokoro_ruslan_b200.features import FeaturePipeline
    from oracle import features as of
    from oracle import melstft as om
    rng = np.random.default_rng(2)
    n_max = 256 * 799 + 100
    lens = [n_max, 150000, 99999, 180224, 64000, n_max, 30000, 120001]
    t = np.arange(n_max) / 22050.0
    wav = np.zeros((8, n_max), np.float32)
    for b, n in enumerate(lens):
        f0 = 90.0 + 40.0 * b + 20.0 * np.sin(2 * np.pi * 0.7 * t)
        ph = 2 * np.pi * np.cumsum(f0) / 22050.0
        x = sum(np.sin(k * ph) / k for k in range(1, 5)) * (0.2 + 0.05 * b) * (np.sin(2 * np.pi * 1.3 * t) > -0.5)
        wav[b, :n] = (x + rng.normal(0, 0.003, n_max))[:n]
    out = FeaturePipeline()(torch.from_numpy(wav).cuda(), torch.tensor(lens))
    assert out["mel_spec"].shape == (8, 80, 800)
    for b in (0, 2, 6):
        n = lens[b]
        x = wav[b, :n] / (np.abs(wav[b, :n]).max() + 1e-9)
        T = 1 + n // 256
        assert int(out["mel_lengths"][b]) == T
        mel = om.log_mel(wav[b, :n])
        assert float(np.abs(out["mel_spec"][b, :, :T].cpu().numpy() - mel).max()) < 1e-3
        p = of.extract_pitch(x)[:T]
        d = np.abs(out["pitch"][b, :T].cpu().numpy() - p)
        assert (d < 1e-4).mean() >= 0.98 and d.max() < 0.1, (b, (d < 1e-4).mean(), d.max())
        e = of.extract_energy_from_mel(np.exp(mel.T.astype(np.float64)).astype(np.float32), False)
        assert float(np.abs(out["energy"][b, :T].cpu().numpy() - e).max()) < 1e-3
        assert float(out["pitch"][b, T:].abs().sum()) == 0.0 and float(out["energy"][b, T:].abs().sum()) == 0.0


def test_resample_matches_torchaudio_fixture():
    from kokoro_ruslan_b200.features import resample, speed_perturb
    f = np.load(os.path.join(HERE, "golden", "resample.npz"))
    x = torch.from_numpy(f["x"]).cuda()
    for factor in f["factors"]:
        new = int(22050 * float(factor))
        want, js, exact = f[f"y_{new}"], f[f"js_{new}"], f[f"exact_{new}"]
        got = resample(x, 22050, new).cpu().numpy()
        assert got.shape == want.shape
        assert np.abs(got[:, js] - exact).max() < 5e-7, factor             # exact application of torchaudio's filter bank
        assert np.abs(got - want).max() < max(2e-6, 2.5 * np.abs(want[:, js] - exact).max()), factor
    y, lens = speed_perturb(x, 0.93, lengths=torch.tensor([12000, 7001]))
    assert lens.tolist() == [11160, 6511] and float(y.abs().amax(dim=1).min()) == pytest.approx(1.0, abs=1e-6)
    assert float(y[1, 6511:].abs().sum()) == 0.0


def test_mel_stft_large_batch_matches_oracle():
    """kr_mel_stft (radix-4 frame kernel) at the bench shape (8 x 800 frames) against the float64 oracle."""
    from kokoro_ruslan_b200.features import LogMelSpectrogram
    from oracle import melstft as om
    tr = LogMelSpectrogram()
    g = torch.Generator().manual_seed(5)
    big = torch.randn(8, 256 * 799 + 100, generator=g)
    a = tr(big.cuda(), peak_normalize=False).cpu()
    assert a.shape == (8, 80, 800)
    want = om.log_mel(big[3].numpy().astype(np.float64), peak_normalize=False)
    assert float(np.abs(a[3].numpy() - want).max()) < 1e-3


def test_trailing_trim_matches_reference_rule():
    from kokoro_ruslan_b200.features import trailing_trim_end
    from test_features_emu_cpu import _reference_trim_end
    g = torch.Generator().manual_seed(0)
    for T, speech_end in ((300, 180), (300, 299), (90, 20), (40, 10), (500, 0)):
        mel = torch.full((T, 80), -10.5) + 0.3 * torch.randn(T, 80, generator=g)
        mel[:speech_end] = -5.0 + 2.0 * torch.randn(speech_end, 80, generator=g)
        mel = mel.clamp(-11.5, 2.0)
        assert int(trailing_trim_end(mel.cuda())) == _reference_trim_end(mel), T
