#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_train_step_gpu.py tests/test_dropout_gpu.py tests/test_attn_gpu.py tests/test_norm_fused_gpu.py tests/test_parity_configs_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_engine.log 2>&1; tail -4 gpurun_out/pytest_engine.log
timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_step.log 2>&1; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_step.log") if l.startswith("{")][-1]); print("STEP", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"].get("frac"), d["gpu_launches"])
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_step.log").read()[-2000:])
PY
