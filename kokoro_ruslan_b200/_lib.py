"""ctypes loader for libkokoro_b200.so — the only gateway to the CUDA kernels.

There is deliberately no fallback: if the shared object is missing or a kernel reports an
error, a RuntimeError is raised (the reference trainer swallows per-batch RuntimeErrors and
continues, reference src/kokoro/training/trainer.py:2679-2686, so errors must be exceptions).
"""
from __future__ import annotations

import ctypes
from pathlib import Path

_LIB = None
LIB_PATH = Path(__file__).resolve().parent / "libkokoro_b200.so"


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m kokoro_ruslan_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot path)")
        _LIB = ctypes.CDLL(str(LIB_PATH))
        _LIB.kr_last_error.restype = ctypes.c_char_p
        _LIB.kr_launch_count.restype = ctypes.c_longlong
    return _LIB


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().kr_last_error().decode(errors="replace")
        raise RuntimeError(f"libkokoro_b200 {what} failed (code {rc}): {msg}")


def launch_count() -> int:
    """Kernels launched by libkokoro_b200 so far (graph replays are not re-counted)."""
    return int(lib().kr_launch_count())
