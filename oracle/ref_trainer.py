"""ORACLE-side harness (test infrastructure, never the product path): the UNMODIFIED reference ``KokoroTrainer`` (from baseline/_ref, the pip-installed copy that also travels
to the GPU box) driving a model through its own ``_setup_model`` / ``_setup_optimizer`` / ``_setup_ema`` /
``_setup_weight_norm_constraints`` / ``train_epoch`` — with the INTEGRATION.md section 1 patch (swap the ``KokoroModel``
symbol the trainer constructs) applied by assignment instead of by editing the file.

The trainer object is built with ``__new__`` plus the attributes its epoch loop reads, the pattern of the reference's own
tests/unit/test_trainer_adaptive_stabilization.py:41-137 (no dataset on disk, no TensorBoard, no profilers).

Users: tests/test_ref_trainer_gpu.py (the reference trainer trains the B200 model) and bench.py's reference arm /
cpu_baseline leg (the reference trainer trains the reference's own model on the host cores).
"""
from __future__ import annotations

import logging
import os
import sys
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF, "kokoro"))


def _import_reference():
    # the repository's own `kokoro` shim (shim/kokoro: the B200 implementation behind the reference's import paths) may
    # already be imported in this process; the harness wants the REFERENCE package
    mod = sys.modules.get("kokoro")
    if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(REF):
        for k in [k for k in sys.modules if k == "kokoro" or k.startswith("kokoro.")]:
            del sys.modules[k]
    if REF in sys.path:
        sys.path.remove(REF)
    sys.path.insert(0, REF)
    import kokoro.training.trainer as rt
    from kokoro.training.config import TrainingConfig
    return rt, TrainingConfig


def build_trainer(model_cls: Optional[Callable], device: torch.device, vocab_size: int, overrides: Dict,
                  quiet: bool = True):
    """model_cls = None keeps the reference's own KokoroModel (harness self-check on the CPU)."""
    rt, TrainingConfig = _import_reference()
    if quiet:
        logging.getLogger("kokoro").setLevel(logging.ERROR)
        logging.getLogger(rt.__name__).setLevel(logging.ERROR)
    tc = TrainingConfig()
    for k, v in overrides.items():
        setattr(tc, k, v)
    tr = rt.KokoroTrainer.__new__(rt.KokoroTrainer)
    tr.config = tc
    tr.device = device
    tr.device_type = device.type
    tr.use_mixed_precision, tr.scaler = False, None
    tr.mixed_precision_dtype = torch.bfloat16
    tr._mps_amp_warned = False
    tr.dataset = SimpleNamespace(phoneme_processor=SimpleNamespace(get_vocab_size=lambda: vocab_size))
    tr.dataloader = [None] * 4                    # len() feeds the EMA half-life rule and the scheduler
    tr.batch_sampler = None
    tr.scheduler_per_batch = False
    tr.enable_adaptive_memory = False
    tr.memory_report_interval = 1000
    tr.current_optimizer_step = 0
    tr.optimizer_steps_completed = 0
    tr.profiler = None
    tr.mixed_precision_stats = {k: 0 for k in ("scale_updates", "scale_decreases", "overflow_count", "successful_steps",
                                               "skipped_steps")}
    noop = lambda *a, **k: None                   # noqa: E731
    tr.writer = SimpleNamespace(add_scalar=noop, add_histogram=noop, add_image=noop, flush=noop)
    tr.log_memory_stats = noop
    tr.clear_device_cache = noop
    tr._step_scheduler_with_warmup = noop
    tr.adaptive_memory_cleanup = lambda *a, **k: {"pressure_level": "low", "cleaned": False}
    tr.interbatch_profiler = SimpleNamespace(reset=noop, start_batch=noop, end_batch=noop, start_data_loading=noop,
                                             end_data_loading=noop, start_forward_pass=noop, end_forward_pass=noop,
                                             start_backward_pass=noop, end_backward_pass=noop,
                                             get_statistics=lambda: {}, print_report=noop)
    saved = rt.KokoroModel
    try:
        if model_cls is not None:
            rt.KokoroModel = model_cls           # the one-line import swap
        tr._setup_model()
    finally:
        rt.KokoroModel = saved
    tr._setup_optimizer()
    tr._setup_ema()
    tr._setup_weight_norm_constraints()
    tr._setup_grad_explosion_tracker()
    return tr


def run_epoch(tr, batches: List[Dict[str, torch.Tensor]]):
    """One reference ``train_epoch`` over the given collated batches.  The reference swallows per-batch exceptions
    (trainer.py:2679-2690): re-raise the first one so a broken model cannot pass as 'skipped batches'."""
    rt, _ = _import_reference()
    errors: List[str] = []

    class Grab(logging.Handler):
        def emit(self, record):
            if record.levelno >= logging.ERROR:
                errors.append(record.getMessage())
    h = Grab()
    lg = logging.getLogger(rt.__name__)
    old = lg.level
    lg.setLevel(logging.ERROR)
    lg.addHandler(h)
    try:
        tr.dataloader = batches
        out = tr.train_epoch(0)
    finally:
        lg.removeHandler(h)
        lg.setLevel(old)
    if errors:
        raise AssertionError("reference train_epoch logged errors: " + " | ".join(errors[:3]))
    return out
