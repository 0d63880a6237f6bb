"""Aggregates an ncu CSV with gpu__time_duration.sum + dram__bytes_{read,write}.sum per kernel function
(second half of the launches = the warm step) into profiles/<out>.json / .txt.
usage: python tools/ncu_traffic.py in.csv out_prefix"""
import csv, json, re, sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = defaultdict(dict)
order = []
for r in rows:
    key = r["ID"]
    if key not in per:
        order.append(key)
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        v = v / 1000.0 if u.startswith("n") else (v if u.startswith("u") else v * 1000.0)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        v *= mult
    per[key][name] = v
    per[key]["kernel"] = r["Kernel Name"]
ids = order[len(order) // 2:]
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in ids:
    d = per[k]
    short = re.sub(r"\(.*", "", d["kernel"]).replace("void ", "").replace("<unnamed>::", "")
    short = re.sub(r"<.*", "", short)
    a = agg[short]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
out = {k: {"launches": v[0], "us": v[1], "dram_read_bytes": v[2], "dram_write_bytes": v[3],
           "traffic_per_launch": (v[2] + v[3]) / v[0]} for k, v in agg.items()}
json.dump(out, open(sys.argv[2] + ".json", "w"), indent=1)
with open(sys.argv[2] + ".txt", "w") as f:
    tot = sum(v["us"] for v in out.values())
    f.write(f"{len(ids)} launches, {tot/1000:.3f} ms serialised (ncu, cold cache); DRAM bytes per kernel function\n")
    for k, v in sorted(out.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"{v['us']:9.1f} us {100*v['us']/tot:5.1f}%  n={v['launches']:4d}  dram R {v['dram_read_bytes']/1e6:9.1f} MB  W {v['dram_write_bytes']/1e6:9.1f} MB  "
                f"({(v['dram_read_bytes']+v['dram_write_bytes'])/max(v['us'],1e-9)/1e3:8.1f} GB/s)  {k}\n")
print(open(sys.argv[2] + ".txt").read()[:3000])
