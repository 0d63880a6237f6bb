"""Per-launch CUDA-event timing of the C-ABI calls (the live measurement behind bench.py's
``roofline`` object and profiles/*.md).

``with OpTimer() as t:`` swaps every public function of :mod:`kokoro_ruslan_b200.ops` for a wrapper
that records a CUDA event pair on the launching stream around the call, together with the
algorithmic FLOPs / bytes derived from the argument shapes.  The GPU is first kept busy with a
spin kernel so that the host runs ahead and the event pairs measure kernel durations rather than
launch gaps.  Only eager (non-graph) execution can be instrumented this way.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, Dict, List, Tuple

import torch

from . import ops


def _nbytes(t) -> int:
    return t.numel() * t.element_size() if isinstance(t, torch.Tensor) else 0


def _gemm_work(args, kwargs) -> Tuple[str, float, float]:
    a, b, out = args[0], args[1], args[2]
    a_mn = kwargs.get("a_mn_major", False)
    batch = a.shape[0] if a.dim() == 3 else 1
    a2 = a[0] if a.dim() == 3 else a
    K, M = (a2.shape if a_mn else a2.shape[::-1])
    N = out.shape[-1]
    flops = 2.0 * M * N * K * batch
    # algorithmic bytes: unique operand elements (an overlapping-row conv view counts its base rows once)
    a_el = M * K if a2.stride(0) >= a2.shape[1] or a_mn else M * a2.stride(0) + K
    byts = 2.0 * batch * (a_el + N * K) + _nbytes(out)
    kind = "wgrad" if kwargs.get("accumulate", False) else ("dgrad" if kwargs.get("b_mn_major", False) else "fwd")
    return f"gemm[{kind}] M{M} N{N} K{K}", flops, byts


def _attn_work(name, args) -> Tuple[str, float, float]:
    q, k = args[0], args[1]
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    causal = bool(args[-2])
    mm = 2.0 * B * H * Sq * Sk * D * (0.5 if causal else 1.0)
    flops = mm * (2 if name == "attn_fwd" else 5)
    byts = sum(_nbytes(t) for t in args if isinstance(t, torch.Tensor))
    return f"{name} B{B} H{H} Sq{Sq} Sk{Sk}{' causal' if causal else ''}", flops, byts


class OpTimer:
    def __init__(self, spin_ms: float = 30.0):
        self.records: List[Tuple[str, float, float, torch.cuda.Event, torch.cuda.Event]] = []
        self._saved: Dict[str, Callable] = {}
        self.spin_ms = spin_ms
        # every GEMM call of the instrumented pass (original function, args, kwargs): lets the caller replay exactly
        # those launches back to back inside a CUDA graph (replay_gemms) — kernel time without event-pair overhead
        self.gemm_calls: List[Tuple[Callable, tuple, dict]] = []

    def _wrap(self, name: str, fn: Callable) -> Callable:
        def wrapper(*args, **kwargs):
            if name == "gemm":
                label, flops, byts = _gemm_work(args, kwargs)
                self.gemm_calls.append((fn, args, kwargs))
            elif name in ("attn_fwd", "attn_bwd"):
                label, flops, byts = _attn_work(name, args)
            else:
                label, flops = name, 0.0
                byts = float(sum(_nbytes(t) for t in list(args) + list(kwargs.values())
                                 if isinstance(t, torch.Tensor)))
                for a in args:
                    if isinstance(a, (list, tuple)):
                        byts += sum(_nbytes(t) for t in a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*args, **kwargs)
            e1.record()
            self.records.append((label, flops, byts, e0, e1))
            return r
        return wrapper

    def region(self, label: str, flops: float = 0.0, byts: float = 0.0):
        timer = self

        class _R:
            def __enter__(self_inner):
                self_inner.e0 = torch.cuda.Event(enable_timing=True)
                self_inner.e0.record()

            def __exit__(self_inner, *exc):
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                timer.records.append((label, flops, byts, self_inner.e0, e1))
        return _R()

    def __enter__(self):
        for name in dir(ops):
            fn = getattr(ops, name)
            if name.startswith("_") or not callable(fn) or getattr(fn, "__module__", None) != ops.__name__:
                continue
            if isinstance(fn, type) or name in ("make_drop_spec", "drop_thr", "drop_keep"):
                continue      # host-side helpers: no launch
            self._saved[name] = fn
            setattr(ops, name, self._wrap(name, fn))
        if self.spin_ms > 0:
            torch.cuda._sleep(int(self.spin_ms * 1.9e6))
        return self

    def __exit__(self, *exc):
        for name, fn in self._saved.items():
            setattr(ops, name, fn)
        torch.cuda.synchronize()

    def replay_gemms(self, reps: int = 5) -> Tuple[float, int]:
        """(ms per pass, launches): all recorded GEMM launches re-issued in order on one stream inside a CUDA graph
        and timed with CUDA events over `reps` replays.  The operands are the step's own tensors (the caller keeps
        them alive); outputs are simply rewritten (weight-gradient GEMMs accumulate once more, which nothing reads
        before the next step zeroes the buffer)."""
        if not self.gemm_calls:
            return 0.0, 0
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for fn, a, kw in self.gemm_calls:          # warm (function attributes, tensor maps)
                fn(*a, **kw)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for fn, a, kw in self.gemm_calls:
                    fn(*a, **kw)
            g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
            for _ in range(reps):
                g.replay()
            e1.record(side)
            side.synchronize()
        torch.cuda.current_stream().wait_stream(side)
        return e0.elapsed_time(e1) / reps, len(self.gemm_calls)

    def summary(self) -> List[dict]:
        """Per-label totals sorted by device time."""
        agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
        for label, flops, byts, e0, e1 in self.records:
            a = agg[label]
            a[0] += 1
            a[1] += e0.elapsed_time(e1)
            a[2] += flops
            a[3] += byts
        total = sum(a[1] for a in agg.values()) or 1.0
        rows = [{"op": k, "launches": v[0], "ms": v[1], "share": v[1] / total, "flops": v[2], "bytes": v[3],
                 "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0,
                 "gbs": v[3] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0} for k, v in agg.items()]
        rows.sort(key=lambda r: -r["ms"])
        return rows
