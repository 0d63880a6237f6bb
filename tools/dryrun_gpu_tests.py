"""Dry run of `-m gpu` test modules WITHOUT a device: libkokoro_b200.so is replaced by a recording stand-in
(tests/test_abi_calls_cpu.py RecordingLib), `.cuda()` / device="cuda" are mapped to the CPU, and every test function is
called directly.  Nothing is computed, so a test is expected to stop at its first DATA-dependent assertion; what the run
flushes out before any GPU minute is spent are Python-level errors on the device path — wrong keyword, wrong constructor
arguments, a ctypes arity mismatch, a shape assert in a wrapper.  (It found the `KokoroModel(vocab_size=59)` constructor
bug of the first inference tests.)

    python tools/dryrun_gpu_tests.py [test_module ...]        # default: the the four late-round-1 modules

Limits: parametrised tests are reported as TypeError (call them by hand), and tests that use torch's OWN CUDA features
(torch.cuda.synchronize, torch's fused AdamW as a checker) cannot be dry-run.
"""
import ctypes
import importlib
import inspect
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
os.environ["KR_STREAMS"] = "0"; os.environ["KR_DECODE_GRAPH"]="0"
import test_abi_calls_cpu as T
from kokoro_ruslan_b200 import _lib, features, ops, engine as E, params, optim, train_step as TS, hifigan as HG, model as M, inference as INF, lengths as LN
lib = T.RecordingLib({"kr_dec_state_size": 256, "kr_val_metrics_acc_floats": 128, "kr_optim_ctrl_size": 64, "kr_resample_length": 11160})
class Finish:
    restype = ctypes.c_int
    def __call__(self, *args):
        st = (ctypes.c_int * 8).from_address(args[0].value)
        if st[1]: return 0
        st[0] += 1
        if st[0] >= st[3] + 2 or st[0] >= st[4]: st[1], st[2] = 1, st[0]      # two frames past lo, or the hard bound hi
        return 0
object.__setattr__(lib, "kr_dec_finish", Finish())
ptr = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
for mod in (ops, features, E, params, optim, TS, HG, M, INF, LN):
    if hasattr(mod, "lib"): setattr(mod, "lib", lambda: lib)
    if hasattr(mod, "_ptr"): setattr(mod, "_ptr", ptr)
    if hasattr(mod, "_stream"): setattr(mod, "_stream", lambda: ctypes.c_void_p(0))
ops._p = lambda t: None if t is None else t.data_ptr()
features._need_cuda = lambda t, w: None
_lib.lib = lambda: lib
torch.cuda.is_available = lambda: True
E.AcousticEngine._empty = lambda self, *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype)   # deterministic dry run
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.cuda.current_stream = lambda *a, **k: None
torch.Tensor.cuda = lambda self, *a, **k: self
torch.Tensor.is_cuda = property(lambda self: True)
_orig_to = torch.Tensor.to
def _to(self, *a, **k):
    a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) or (isinstance(x, torch.device) and x.type == "cuda") else x for x in a)
    if "device" in k and str(k["device"]).startswith("cuda"): k["device"] = "cpu"
    return _orig_to(self, *a, **k)
torch.Tensor.to = _to
_orig_empty, _orig_zeros, _orig_full, _orig_tensor, _orig_arange = torch.empty, torch.zeros, torch.full, torch.tensor, torch.arange
def _fix(fn):
    def g(*a, **k):
        if "device" in k and str(k["device"]).startswith("cuda"): k["device"] = "cpu"
        return fn(*a, **k)
    return g
for n in ("empty","zeros","full","tensor","arange","ones","empty_like","linspace","randn"):
    setattr(torch, n, _fix(getattr(torch, n)))
_orig_device = torch.device
import pytest
class MP:
    def setenv(self, k, v): os.environ[k] = v
MODULES = sys.argv[1:] or ["test_features_gpu", "test_metrics_gpu", "test_average_by_duration_gpu", "test_inference_gpu"]
bad = 0
for modname in MODULES:
    mod = importlib.import_module(modname)
    for name in [n for n in dir(mod) if n.startswith("test_")]:
        fn = getattr(mod, name)
        try:
            fn(MP()) if "monkeypatch" in inspect.signature(fn).parameters else fn()
            print(modname, name, "-> ran to the end")
        except AssertionError as e:
            tb = traceback.extract_tb(e.__traceback__)[-1]
            print(modname, name, "-> AssertionError at", os.path.basename(tb.filename), tb.lineno, "(numeric check; expected without a device)")
        except Exception as e:
            tb = traceback.extract_tb(e.__traceback__)
            bad += 1
            print(modname, name, "-> ", type(e).__name__, str(e)[:150], "at", [(os.path.basename(t.filename), t.lineno) for t in tb[-3:]])
print(f"{bad} test(s) failed before reaching a data-dependent assertion")
sys.exit(1 if bad else 0)
