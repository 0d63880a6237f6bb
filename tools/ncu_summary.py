"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
LAST step (second half of the launches).  usage: python tools/ncu_summary.py file.csv [skip_frac]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * skip):]
agg = defaultdict(lambda: [0, 0.0])
for name, us, grid, block in rows:
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"void ", "", short)
    short = short[:110]
    agg[short][0] += 1
    agg[short][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot/1000:.3f} ms total (serialised, cold cache)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {100*v[1]/tot:5.1f}%  n={v[0]:4d}  avg {v[1]/v[0]:7.1f} us  {k}")
