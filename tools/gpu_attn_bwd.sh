#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_dropout_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; tail -5 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log

