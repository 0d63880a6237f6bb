#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_dropout_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; tail -5 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_train_step_gpu.py tests/test_ref_trainer_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_engine.log 2>&1; tail -8 gpurun_out/pytest_engine.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_attn.log 2>&1; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_attn.log") if l.startswith("{")][-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"].get("frac"))
    for r in d["roofline"]["top_ops"]: print(r)
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_attn.log").read()[-2000:])
PY
