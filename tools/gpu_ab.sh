#!/bin/bash
# A/B of env-selected variants of the step: usage gpu_ab.sh "VAR=a" "VAR=b" ...
mkdir -p gpurun_out
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_ab.log 2>&1
  python - "$v" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open("gpurun_out/bench_ab.log") if l.startswith("{")][-1]); print(sys.argv[1], "STEP", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],4), d["gpu_launches"])
except Exception as e: print(sys.argv[1], "ERR", e); print(open("gpurun_out/bench_ab.log").read()[-1500:])
PY
done
