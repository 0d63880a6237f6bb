"""Host emulation of kr_val_metrics (csrc/kr_metrics_core.cuh compiled by g++ -DKR_HOST_EMU) against a literal
restatement of the reference's validation-metric loop (training/trainer.py:1868-1916)."""
import ctypes
import math
import os
import subprocess

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=["seq", "simt"])
def emu(request, tmp_path_factory):
    """The kernels' bodies as a host library: "seq" = one sequential thread per block; "simt" = one host thread per CUDA
    thread with the real block / warp geometry, barriers and warp reductions (tests/emu/emu_simt.h)."""
    simt = request.param == "simt"
    so = tmp_path_factory.mktemp("emu") / ("metrics_emu_%s.so" % request.param)
    flags = ["-DKR_HOST_EMU_SIMT", "-pthread", "-I", os.path.join(HERE, "emu")] if simt else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-x", "c++", "-I",
                    os.path.join(ROOT, "kokoro_ruslan_b200", "csrc"), os.path.join(HERE, "emu", "metrics_emu.cpp"), "-o", str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    lib.simt = simt
    return lib


def reference_loop(batches):
    """trainer.py:1868-1916 verbatim in structure: per-sample loop, batch means, epoch means of the batch means."""
    sc_sum = f0_sum = 0.0
    sc_n = f0_n = 0
    for mel_pred, mel, pitch_pred, pitch, lengths in batches:
        bs, bc = 0.0, 0
        for b in range(mel.size(0)):
            L = int(lengths[b])
            if L <= 0:
                continue
            num, den = torch.norm(mel[b, :L] - mel_pred[b, :L], p="fro"), torch.norm(mel[b, :L], p="fro")
            if den.item() > 0:
                bs += (num / den).item()
                bc += 1
        if bc:
            sc_sum += bs / bc
            sc_n += 1
        bf, bn = 0.0, 0
        for b in range(pitch.size(0)):
            L = int(lengths[b])
            if L <= 0:
                continue
            bf += math.sqrt(torch.mean((pitch[b, :L] - pitch_pred[b, :L]) ** 2).item())
            bn += 1
        if bn:
            f0_sum += bf / bn
            f0_n += 1
    return sc_sum / sc_n, f0_sum / f0_n


def test_metrics_match_the_reference_loop(emu):
    g = torch.Generator().manual_seed(0)
    batches = []
    for B, T, lens in ((4, 50, [50, 31, 0, 7]), (1, 20, [20]), (3, 64, [64, 64, 1])):
        mel = torch.randn(B, T, 80, generator=g) * 2 - 5
        mel_pred = mel + 0.3 * torch.randn(B, T, 80, generator=g)
        pitch, pitch_pred = torch.rand(B, T, generator=g), torch.rand(B, T, generator=g)
        batches.append((mel_pred, mel, pitch_pred, pitch, torch.tensor(lens, dtype=torch.int64)))
    n = emu.emu_val_metrics_acc_floats()
    acc = torch.zeros(n)
    p = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
    for mel_pred, mel, pitch_pred, pitch, lens in batches:
        B, T, C = mel.shape
        assert emu.emu_val_metrics(p(mel_pred), p(mel), p(pitch_pred), p(pitch), p(lens), p(acc), B, T, T, C) == 0
    want_sc, want_f0 = reference_loop(batches)
    assert acc[1] == 3 and acc[3] == 3
    assert float(acc[0] / acc[1]) == pytest.approx(want_sc, rel=1e-5)
    assert float(acc[2] / acc[3]) == pytest.approx(want_f0, rel=1e-5)
    assert acc[4:8].abs().sum() == 0                 # the arrival counter is back at zero for the next batch
