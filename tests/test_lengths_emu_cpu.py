"""Host emulation of kr_average_by_duration (csrc/kr_lengths_core.cuh compiled by g++ -DKR_HOST_EMU) against fixtures
from the LIVE reference function (tests/golden/make_golden_average.py) and the reference's own unit-test values
(tests/unit/test_utils_lengths.py:77-91): bit-identical, quirks included."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=["seq", "simt"])
def emu(request, tmp_path_factory):
    """The kernels' bodies as a host library: "seq" = one sequential thread per block; "simt" = one host thread per CUDA
    thread with the real block / warp geometry, barriers and warp reductions (tests/emu/emu_simt.h)."""
    simt = request.param == "simt"
    so = tmp_path_factory.mktemp("emu") / ("lengths_emu_%s.so" % request.param)
    flags = ["-DKR_HOST_EMU_SIMT", "-pthread", "-I", os.path.join(HERE, "emu")] if simt else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-x", "c++", "-I",
                    os.path.join(ROOT, "kokoro_ruslan_b200", "csrc"), os.path.join(HERE, "emu", "lengths_emu.cpp"), "-o", str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    lib.simt = simt
    return lib


def run(lib, values, dur, mask=None):
    values, dur = values.float().contiguous(), dur.long().contiguous()
    B, T = values.shape
    P = dur.shape[1]
    label, out = torch.zeros(B, T, dtype=torch.int32), torch.full((B, P), float("nan"))
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())        # noqa: E731
    assert lib.emu_average_by_duration(p(values), p(dur), p(m), p(label), p(out), B, P, T) == 0
    return out


def test_reference_unit_test_values(emu):
    values = torch.tensor([[1.0, 2.0, 3.0, 4.0], [10.0, 20.0, 30.0, 40.0]])
    durations = torch.tensor([[1, 3], [0, 4]])
    avg = run(emu, values, durations)
    assert torch.allclose(avg[0], torch.tensor([1.0, 3.0])) and torch.allclose(avg[1], torch.tensor([0.0, 25.0]))
    avg2 = run(emu, values, durations, torch.tensor([[False, True], [False, False]]))
    assert torch.allclose(avg2[0], torch.tensor([1.0, 0.0]))


def test_bit_identical_to_live_reference_fixtures(emu):
    f = np.load(os.path.join(HERE, "golden", "average_by_duration.npz"))
    for k in range(4):
        v, d, m = (torch.from_numpy(f[f"{n}{k}"]) for n in "vdm")
        assert torch.equal(run(emu, v, d), torch.from_numpy(f[f"a{k}"])), k
        assert torch.equal(run(emu, v, d, m), torch.from_numpy(f[f"am{k}"])), k


def test_reference_quirks_are_reproduced(emu):
    """Values checked against the live reference function when this test was written."""
    # frames 3..5 are covered by no token: their label is 0, so token 0 averages them too -> (1+2+4+5+6)/5, not 1.5
    got = run(emu, torch.tensor([[1.0, 2, 3, 4, 5, 6]]), torch.tensor([[2, 1]]))
    assert torch.allclose(got, torch.tensor([[3.6, 3.0]]))
    # durations past the last frame: tokens 1 and 2 pile their labels (1 + 2 = 3 >= P) onto frame T-1, which drops out
    got = run(emu, torch.tensor([[1.0, 2, 3, 4]]), torch.tensor([[3, 2, 5]]))
    assert torch.allclose(got, torch.tensor([[2.0, 0.0, 0.0]]))


def test_kernel_source_equals_the_installed_reference_on_edge_durations(emu):
    """kr_average_by_duration's own source side by side with the installed reference function (utils/lengths.py:156-208) on
    zero / negative / single / huge durations, durations running past the frame axis, and masked phonemes: bit-identical."""
    import logging
    import sys
    sys.path.insert(0, ROOT)
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.utils.lengths import average_by_duration
    g = torch.Generator().manual_seed(9)
    cases = {"all zero": torch.zeros(2, 5, dtype=torch.long), "one row zero": torch.tensor([[0, 0, 0, 0], [3, 0, 2, 1]]),
             "single token": torch.tensor([[7]]), "negative durations": torch.tensor([[2, -3, 4], [-1, -1, 5]]),
             "ragged": torch.randint(0, 9, (3, 11), generator=g), "long utterance": torch.randint(1, 16, (2, 260), generator=g),
             "one huge duration": torch.tensor([[1, 1500, 1], [2, 2, 2]])}
    for label, dur in cases.items():
        B, P = dur.shape
        total = int(dur.clamp(min=0).sum(1).max())
        for T in sorted({max(1, total - 2), max(1, total), total + 3}):     # frame axis shorter than / equal to / longer than the durations
            vals = torch.rand(B, T, generator=g)
            for mask in (None, torch.rand(B, P, generator=g) > 0.3):
                want = average_by_duration(vals, dur, mask)
                got = run(emu, vals, dur, mask)
                assert got.shape == want.shape and torch.equal(got, want), (label, T, mask is None, got, want)
