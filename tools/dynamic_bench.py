"""BASELINE.json configs[1] as a throughput number: the training step over DYNAMIC batches (DynamicFrameBatchSampler,
max_frames = 8000, ragged utterances of 200 - 1200 frames), the way `kokoro-train` feeds it.  Almost every batch of an epoch has
its own (B, P, T, T') shape and the sampler re-packs its buckets every epoch, so the CUDA-graph cache (one graph per shape, LRU)
rarely hits: this measures the EAGER launch path (same kernels, launched one by one through the C ABI) against the
fixed-shape graph replay of bench.py.  Prints real (un-padded) and padded mel frames per second.
usage: python tools/dynamic_bench.py [n_utterances]"""
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokoro_ruslan_b200.cli import SyntheticDataset  # noqa: E402
from kokoro_ruslan_b200.data import DynamicFrameBatchSampler, collate_fn  # noqa: E402
from kokoro_ruslan_b200.engine import DropoutConfig  # noqa: E402
from kokoro_ruslan_b200.params import ModelConfig  # noqa: E402
from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep  # noqa: E402


def epoch_batches(ds, seed):
    random.seed(seed)
    torch.manual_seed(seed)
    sampler = DynamicFrameBatchSampler(ds, max_frames=8000, min_batch_size=4, max_batch_size=32, shuffle=True)
    return [collate_fn([ds[i] for i in idx], pin_memory=True) for idx in sampler]


PROF = {}


def _timed(obj, name):
    fn = getattr(obj, name)

    def wrap(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            d = PROF.setdefault(name, [0, 0.0])
            d[0] += 1
            d[1] += time.perf_counter() - t
    setattr(obj, name, wrap)


def run(ts, batches, host_api):
    torch.cuda.synchronize()
    PROF.clear()
    t0 = time.perf_counter()
    for b in batches:
        if host_api:
            ts.train_step_host(b)
        else:
            ts.train_step(b)
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 320
    ds = SyntheticDataset(n, min_frames=200, max_frames=1200)
    cfg = ModelConfig()
    for graphs in ((False, True) if os.environ.get("KR_DYN_EAGER_FIRST") else (True, False)):
        ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=100000), device="cuda:0", use_graphs=graphs,
                       dropout=DropoutConfig.reference_training())
        ts.store.init_default(seed=0)
        if os.environ.get("KR_DYN_PROFILE"):
            for name in ("stage", "_new_staged", "_run_fwd_bwd", "_run_optimizer", "_fwd_bwd"):
                _timed(ts, name)
            _timed(ts.opt, "set_lrs")
            _timed(ts.engine, "forward")
            _timed(ts.engine, "losses")
            _timed(ts.engine, "_geom") if hasattr(ts.engine, "_geom") else None
        run(ts, epoch_batches(ds, 0), False)                       # warm-up epoch (lazy kernel attributes, allocator)
        for ep, host_api in ((1, False), (2, True)):
            batches = epoch_batches(ds, ep)
            real = sum(int(b["mel_lengths"].sum()) for b in batches)
            padded = sum(b["mel_specs"].shape[0] * b["mel_specs"].shape[1] for b in batches)
            shapes = len({(b["mel_specs"].shape[0], b["phoneme_indices"].shape[1], b["mel_specs"].shape[1]) for b in batches})
            s = run(ts, batches, host_api)
            print(f"graphs={graphs} epoch {ep} ({'train_step_host' if host_api else 'train_step'}): {len(batches)} batches, "
                  f"{shapes} distinct shapes, {s / len(batches) * 1e3:.2f} ms per step, {real / s / 1e3:.0f} k real mel-frames/s "
                  f"({padded / s / 1e3:.0f} k padded), {len(ts._staged)} shapes cached", flush=True)
            if PROF:
                print("   host time per call (ms): " + ", ".join(f"{k} {v[1] / max(1, v[0]) * 1e3:.2f} x{v[0]}" for k, v in PROF.items()),
                      f"| graphs captured: {sum(g is not None for st in ts._staged.values() for g in st.graph)}", flush=True)
        del ts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
