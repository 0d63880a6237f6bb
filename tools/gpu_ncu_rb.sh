#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"hifi_resblock|kr_gemm" --csv --log-file gpurun_out/rb_launches.csv python tools/hifigan_one.py > /dev/null 2>&1
python - <<'PY'
import csv,re
from collections import defaultdict
lines=[l for l in open('gpurun_out/rb_launches.csv') if l.startswith('"')]
rows=list(csv.DictReader(lines)); per=defaultdict(dict); order=[]
for r in rows:
    k=r["ID"]
    if k not in per: order.append(k)
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]; n=r["Metric Name"]
    if n=="gpu__time_duration.sum": v = v/1000.0 if u.startswith("n") else v
    else: v*={"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
    per[k][n]=v; per[k]["kernel"]=r["Kernel Name"]
ids=order[-68:]
for i,k in enumerate(ids):
    x=per[k]; t=x["gpu__time_duration.sum"]; rd=x.get("dram__bytes_read.sum",0); wr=x.get("dram__bytes_write.sum",0)
    nm="RB" if "resblock" in x["kernel"] else re.sub(r".*kr_gemm_kernel<([^>]*)>.*",r"\1",x["kernel"])
    print(f"{i:3d} {nm:20s} {t:8.1f} us R {rd/1e6:7.1f} W {wr/1e6:7.1f} MB {(rd+wr)/t/1e3:7.1f} GB/s")
PY
