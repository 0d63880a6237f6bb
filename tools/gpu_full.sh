#!/bin/bash
# Full GPU validation: the -m gpu suite, smoke(), the default bench and the reference arm.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rfE -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -30
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; tail -c 2500 gpurun_out/bench_n1.log; tail -5 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n1.log") if l.startswith("{")][-1])
print("STEP", d["ms_per_step"], "e2e", d["e2e"], "frac", d["roofline"].get("frac"), "launches", d["gpu_launches"])
print("HIFI", {k: v for k, v in d["roofline"].get("hifigan", {}).items() if not isinstance(v, (list, dict))})
print("CPU", d.get("cpu_baseline"))
PY
