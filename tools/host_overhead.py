"""Where the HOST time of an eager training step goes (the path dynamic batching takes: tools/dynamic_bench.py).
Runs TrainStep on the CPU against a stub library whose entry points return at once (the dry-run set-up of
tests/test_abi_calls_cpu.py) and profiles the Python side: wrappers, argument marshalling, tensor bookkeeping.  The C side
of a launch (tensor-map encoding, cudaLaunchKernelEx) is NOT in these numbers.
usage: python tools/host_overhead.py [steps]"""
import cProfile
import ctypes
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokoro_ruslan_b200 import _lib, engine as engine_mod, features, ops  # noqa: E402


class StubLib:
    def __init__(self):
        self.n = 0
        self.ret = {"kr_dec_state_size": 256, "kr_val_metrics_acc_floats": 128, "kr_optim_ctrl_size": 64}

    def __getattr__(self, name):
        rv = self.ret.get(name, 0)

        def fn(*args):
            self.n += 1
            return rv
        fn.restype = ctypes.c_int
        object.__setattr__(self, name, fn)
        return fn


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    os.environ["KR_STREAMS"] = "0"             # one stream: there is no device
    lib = StubLib()
    for mod in (ops, features):
        mod.lib = lambda: lib
        mod._ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    ops._p = lambda t: None if t is None else t.data_ptr()
    _lib.lib = lambda: lib
    torch.cuda.is_available = lambda: True
    torch.cuda.current_stream = lambda *a, **k: type("S", (), {"cuda_stream": 0, "wait_stream": lambda s, o: None,
                                                                 "wait_event": lambda s, e: None})()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    engine_mod.AcousticEngine._empty = lambda self, *shape, dtype=torch.float32: torch.empty(*shape, dtype=dtype)
    from bench import synthetic_batch
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    from kokoro_ruslan_b200.engine import DropoutConfig
    cfg = ModelConfig()
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=1000), device="cpu", use_graphs=False,
                   dropout=DropoutConfig.reference_training())
    batch = synthetic_batch(4, 32, 200, cfg.mel_dim, cfg.vocab_size, 1)
    ts.train_step(batch)
    t0 = time.perf_counter()
    for _ in range(steps):
        ts.train_step(batch)
    print(f"{(time.perf_counter() - t0) / steps * 1e3:.2f} ms of Python per step (no profiler)")
    n0, t0 = lib.n, time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        ts.train_step(batch)
    pr.disable()
    dt = (time.perf_counter() - t0) / steps
    calls = (lib.n - n0) / steps
    print(f"{calls:.0f} library calls per step, {dt * 1e3:.2f} ms of Python per step under cProfile")
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
