#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rfE -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -30
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.log
KR_PDL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_pdl.log 2>&1; python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.log","gpurun_out/bench_pdl.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["e2e"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
