#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rfE -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/hifigan_bench.py > gpurun_out/hifigan_bench.log 2>&1; tail -8 gpurun_out/hifigan_bench.log
