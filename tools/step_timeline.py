"""Multi-stream timeline of the graph-replayed training step (bench shape) from CUPTI kernel records (torch.profiler):
per-stream busy time, the idle gaps of the main stream (what it waits for), and what runs at the very end of the step.
usage: python tools/step_timeline.py [out.json]"""
import json
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import B_PER_GPU, N_MELS, P_LEN, T_LEN, synthetic_batch  # noqa: E402
from kokoro_ruslan_b200.engine import DropoutConfig  # noqa: E402
from kokoro_ruslan_b200.params import ModelConfig  # noqa: E402
from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep  # noqa: E402

cfg = ModelConfig()
ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=1000), device="cuda:0", use_graphs=True,
               dropout=DropoutConfig.reference_training())
ts.store.init_default(seed=0)
host = {k: v.pin_memory() for k, v in synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, cfg.vocab_size, 1).items()}
for _ in range(5):
    ts.train_step(host)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        ts.train_step(host)
    torch.cuda.synchronize()
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "step_trace.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
print(len(ev), "GPU activities in 3 steps")
# split into steps at the largest gaps between consecutive activity END and next START across all streams
ends = 0.0
bounds = []
for i, e in enumerate(ev):
    if i and e["ts"] - ends > 30.0:
        bounds.append(i)
    ends = max(ends, e["ts"] + e["dur"])
print("idle gaps > 30 us across ALL streams at activity indices", bounds[:12])
# take the middle step: from the H2D copy burst before it to the next one
h2d = [i for i, e in enumerate(ev) if e.get("cat") == "gpu_memcpy" and "HtoD" in e["name"]]
starts = [h2d[0]] + [h2d[i] for i in range(1, len(h2d)) if ev[h2d[i]]["ts"] - ev[h2d[i - 1]]["ts"] > 1000.0]
print("step starts at activity", starts)
a, b = (starts[1], starts[2]) if len(starts) >= 3 else (0, len(ev))
step = ev[a:b]
t0 = step[0]["ts"]
t1 = max(e["ts"] + e["dur"] for e in step)
print(f"step span {(t1 - t0) / 1e3:.3f} ms, {len(step)} activities")
by = defaultdict(list)
for e in step:
    by[e["args"].get("stream", -1)].append(e)
main = max(by, key=lambda s: len(by[s]))
for s, es in sorted(by.items(), key=lambda kv: -len(kv[1])):
    busy = sum(e["dur"] for e in es)
    print(f"  stream {s}: {len(es):4d} activities, busy {busy / 1e3:6.3f} ms, first at {(es[0]['ts'] - t0) / 1e3:6.3f}, last ends {(es[-1]['ts'] + es[-1]['dur'] - t0) / 1e3:6.3f} ms"
          + ("   <- main" if s == main else ""))
# union busy time (any stream) and concurrency histogram
pts = sorted([(e["ts"], 1) for e in step] + [(e["ts"] + e["dur"], -1) for e in step])
lvl, last, hist = 0, t0, defaultdict(float)
for t, d in pts:
    hist[lvl] += t - last
    lvl += d
    last = t
print("  time with k kernels in flight:", {k: round(v / 1e3, 3) for k, v in sorted(hist.items())})
es = by[main]
gaps = []
for p, n in zip(es, es[1:]):
    g = n["ts"] - (p["ts"] + p["dur"])
    if g > 4.0:
        gaps.append((g, p, n))
gaps.sort(key=lambda x: -x[0])
print(f"  main stream: {sum(g for g, _, _ in gaps) / 1e3:.3f} ms in {len(gaps)} gaps > 4 us; the largest:")
for g, p, n in gaps[:14]:
    others = [o for o in step if o["args"].get("stream") != main and o["ts"] < n["ts"] and o["ts"] + o["dur"] > p["ts"] + p["dur"]]
    print(f"    {g:7.1f} us at {(p['ts'] + p['dur'] - t0) / 1e3:6.3f} ms after {p['name'][:48]:48s} before {n['name'][:40]:40s} | meanwhile: "
          + ", ".join(sorted({o['name'][:28] for o in others})[:4]))
print("  last 8 activities of the step:")
for e in sorted(step, key=lambda e: e["ts"] + e["dur"])[-8:]:
    print(f"    ends {(e['ts'] + e['dur'] - t0) / 1e3:6.3f} ms  dur {e['dur']:7.1f} us  stream {e['args'].get('stream')}  {e['name'][:70]}")
# the forward chain: the side stream that carries the most busy time before the main stream resumes
if gaps:
    g0, p0, n0 = gaps[0]
    fwd_end = n0["ts"]
    cand = {s: sum(e["dur"] for e in es_ if e["ts"] < fwd_end) for s, es_ in by.items() if s != main}
    fs = max(cand, key=cand.get)
    chain = [e for e in by[fs] if e["ts"] < fwd_end]
    print(f"  forward chain on stream {fs}: {len(chain)} kernels, busy {sum(e['dur'] for e in chain) / 1e3:.3f} ms, "
          f"span {(chain[-1]['ts'] + chain[-1]['dur'] - chain[0]['ts']) / 1e3:.3f} ms")
    fam = defaultdict(lambda: [0, 0.0, 0.0])
    prev_end = chain[0]["ts"]
    for e in chain:
        k = e["name"].replace("(anonymous namespace)::", "").replace("void ", "")[:44]
        fam[k][0] += 1
        fam[k][1] += e["dur"]
        fam[k][2] += max(0.0, e["ts"] - prev_end)
        prev_end = e["ts"] + e["dur"]
    print("    kernel family                                   n   busy us  gap-before us")
    for k, (n, d, g) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"    {k:44s} {n:4d} {d:9.1f} {g:9.1f}")
    if os.environ.get("KR_TIMELINE_CHAIN"):
        prev_end = chain[0]["ts"]
        for e in chain:
            print(f"      +{(e['ts'] - t0) / 1e3:6.3f} ms gap {e['ts'] - prev_end:5.1f} dur {e['dur']:6.1f}  "
                  f"{e['name'].replace('(anonymous namespace)::', '')[:60]} grid {e['args'].get('grid')}")
            prev_end = e["ts"] + e["dur"]
