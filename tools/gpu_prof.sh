#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_ops.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_ops.log") if l.startswith("{")][-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"].get("frac"), d["launches_per_step"])
for r in d["roofline"]["top_ops"]: print(r)
PY
bash tools/ncu_lists.sh
python tools/ncu_traffic.py gpurun_out/step_launches.csv gpurun_out/r02_step_traffic_v1 && head -50 gpurun_out/r02_step_traffic_v1.txt
