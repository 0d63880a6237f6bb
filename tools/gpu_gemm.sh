#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/gemm_bench.py static > gpurun_out/gemm_static.log 2>&1; cat gpurun_out/gemm_static.log
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_engine_gpu.py tests/test_train_step_gpu.py tests/test_hifigan_gpu.py tests/test_model_gpu.py tests/test_inference_gpu.py tests/test_optim_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; tail -6 gpurun_out/pytest_gemm.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_static.log 2>&1
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_static.log") if l.startswith("{")][-1]); print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"].get("frac"), d["roofline"]["hifigan"].get("ms_per_batch"))
PY
