"""ThreadSanitizer over the SIMT emulation of the dual-compiled kernel bodies (tests/emu/emu_simt.h, tests/emu/tsan_main.cpp):
every body runs with its real block geometry on host threads whose only synchronisation are the barriers the kernel source
asks for — a hand-off through shared memory (or a global scratch row) without a __syncthreads() is a data race TSan
reports.  The negative control removes one barrier from a copy of the source and must be caught.  Host-side counterpart of
`compute-sanitizer --tool racecheck` for the kernels that have not had a hardware run yet."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "kokoro_ruslan_b200", "csrc")
ENV = {**os.environ, "TSAN_OPTIONS": "halt_on_error=1 exitcode=66"}


def _build(include_dir, out):
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I", include_dir, "-I",
                        os.path.join(HERE, "emu"), os.path.join(HERE, "emu", "tsan_main.cpp"), "-o", str(out)],
                       capture_output=True, text=True)
    if r.returncode != 0 and ("tsan" in r.stderr.lower() or "sanitize" in r.stderr.lower()):
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]


def test_kernel_bodies_are_race_free_under_the_simt_emulation(tmp_path):
    exe = tmp_path / "tsan_main"
    _build(CSRC, exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-3000:])
    assert "simt emulation done" in r.stdout and "ThreadSanitizer" not in r.stderr


def test_a_removed_barrier_is_reported(tmp_path):
    """Negative control: without the barrier between the KV-cache append and the key loop of dec_attn_body the run must
    fail with TSan's exit code."""
    broken = tmp_path / "csrc"
    broken.mkdir()
    for f in os.listdir(CSRC):
        if f.endswith(".cuh"):
            shutil.copy(os.path.join(CSRC, f), broken / f)
    p = broken / "kr_decode_core.cuh"
    s = p.read_text()
    needle = "  KRD_SYNC();\n  // phase B: ONE KEY PER LANE"
    assert needle in s
    p.write_text(s.replace(needle, "  // phase B: ONE KEY PER LANE", 1))
    exe = tmp_path / "tsan_broken"
    _build(str(broken), exe)
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=ENV, timeout=300)
    assert r.returncode == 66 and "data race" in r.stderr and "dec_attn_body" in r.stderr
