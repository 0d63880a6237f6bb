"""Host emulation of kr_resample (csrc/kr_resample_core.cuh compiled by g++ -DKR_HOST_EMU) against fixtures from the
LIVE torchaudio (tests/golden/make_golden_resample.py): torchaudio.functional.resample's output and the exact
(float64-accumulated) application of torchaudio's own filter bank."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SR = 22050


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "resample_emu.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-I", os.path.join(ROOT, "kokoro_ruslan_b200", "csrc"),
                    os.path.join(HERE, "emu", "resample_emu.cpp"), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    lib.emu_resample_length.restype = ctypes.c_longlong
    return lib


def emu_resample(lib, x, new, lengths=None):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    B, n = x.shape
    m = lib.emu_resample_length(ctypes.c_longlong(n), SR, new)
    y = np.full((B, m), np.nan, np.float32)
    lens = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int64)
    p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)      # noqa: E731
    assert lib.emu_resample(p(x), p(lens), p(y), B, ctypes.c_longlong(n), ctypes.c_longlong(m), SR, new, 6,
                            ctypes.c_float(0.99)) == 0
    return y


def test_matches_torchaudio_filter_bank_and_output(emu):
    f = np.load(os.path.join(HERE, "golden", "resample.npz"))
    for factor in f["factors"]:
        new = int(SR * float(factor))
        want, js, exact = f[f"y_{new}"], f[f"js_{new}"], f[f"exact_{new}"]
        got = emu_resample(emu, f["x"], new)
        assert got.shape == want.shape, (factor, got.shape, want.shape)
        # the algorithm's exact value (torchaudio's own taps, float64 accumulation): float32 rounding only
        assert np.abs(got[:, js] - exact).max() < 2e-7, factor
        # torchaudio's float32 strided conv1d carries up to 4e-4 of its own accumulation noise for ratios that do not
        # reduce (11 k-tap rows); we must be at least as close to the exact value as it is, and within that noise of it
        ref_noise = np.abs(want[:, js] - exact).max()
        assert np.abs(got - want).max() < max(2e-6, 2.5 * ref_noise), (factor, np.abs(got - want).max(), ref_noise)


def test_ragged_rows_equal_single_runs(emu):
    f = np.load(os.path.join(HERE, "golden", "resample.npz"))
    x = f["x"].copy()
    lens = [12000, 7001]
    x[1, 7001:] = 9.0                                        # garbage beyond the length must not leak in
    new = int(SR * 0.93)
    got = emu_resample(emu, x, new, lens)
    single = emu_resample(emu, f["x"][1, :7001], new)[0]
    assert np.array_equal(got[1, :len(single)], single) and np.all(got[1, len(single):] == 0.0)
    assert np.array_equal(got[0], emu_resample(emu, f["x"][0], new)[0])
