#!/bin/bash
# Serialised launch list of one eager training step with DRAM bytes per kernel (profiles/r02_step_traffic_*).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/step_traffic.csv python tools/one_step.py 2 > gpurun_out/step_traffic.log 2>&1
python tools/ncu_traffic.py gpurun_out/step_traffic.csv gpurun_out/r02_step_traffic_v3 | head -30
python - <<'PY'
import csv
from collections import defaultdict
lines=[l for l in open('gpurun_out/step_traffic.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines)); per=defaultdict(dict); order=[]
for r in rows:
    k=r["ID"]
    if k not in per: order.append(k)
    if r["Metric Name"]=="gpu__time_duration.sum":
        v=float(r["Metric Value"].replace(",","")); per[k]["t"]= v/1000.0 if r["Metric Unit"].startswith("n") else v
    per[k]["kernel"]=r["Kernel Name"]
ids=order[len(order)//2:]
print("attn_bwd launches (us):", [round(per[k]["t"],1) for k in ids if "attn_bwd_kernel" in per[k]["kernel"]])
print("attn_fwd launches (us):", [round(per[k]["t"],1) for k in ids if "attn_fwd_kernel" in per[k]["kernel"]])
PY
