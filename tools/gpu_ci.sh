#!/bin/bash
# Runs on the GPU box (via gpurun): GPU test-suite, smoke, bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.log; tail -15 gpurun_out/bench.err
