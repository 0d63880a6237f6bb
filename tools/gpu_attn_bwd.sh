#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attn_gpu.py tests/test_dropout_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; tail -5 gpurun_out/pytest_attn.log
timeout 200 python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 2 -c 1 -o gpurun_out/r02_attn_bwd -f python tools/attn_one.py > gpurun_out/ncu_attn_bwd.log 2>&1; tail -2 gpurun_out/ncu_attn_bwd.log
