#!/bin/bash
# First GPU call of the next round: the kernels written after round 1's GPU budget was spent (features N1, decode N2,
# validation metrics N3) get their first hardware run, then their micro-benchmarks.  Outputs land in gpurun_out/.
#   gpurun --timeout 900 -- 'bash tools/gpu_next.sh'
mkdir -p gpurun_out
python -m pytest tests/test_zz_features_gpu.py tests/test_zz_lengths_gpu.py tests/test_zz_metrics_gpu.py tests/test_zz_inference_gpu.py -m gpu -q -rxX \
  > gpurun_out/pytest_zz.log 2>&1
tail -15 gpurun_out/pytest_zz.log
timeout 120 compute-sanitizer --tool racecheck python -m pytest tests/test_zz_features_gpu.py -m gpu -q -x -k "pitch_matches or energy" \
  > gpurun_out/racecheck_features.log 2>&1; tail -3 gpurun_out/racecheck_features.log
timeout 200 compute-sanitizer --tool racecheck python -m pytest tests/test_zz_inference_gpu.py -m gpu -q -x -k teacher \
  > gpurun_out/racecheck_decode.log 2>&1; tail -3 gpurun_out/racecheck_decode.log
timeout 120 python tools/features_bench.py > gpurun_out/features_bench.log 2>&1; cat gpurun_out/features_bench.log
timeout 200 python tools/decode_bench.py 1 64 400 > gpurun_out/decode_bench.log 2>&1; cat gpurun_out/decode_bench.log
KR_ATTN_FAST=1 python -m pytest tests -m gpu -q -x --deselect tests/test_zz_features_gpu.py --deselect tests/test_zz_inference_gpu.py \
  --deselect tests/test_zz_metrics_gpu.py --deselect tests/test_zz_lengths_gpu.py > gpurun_out/pytest_attn_fast.log 2>&1; tail -3 gpurun_out/pytest_attn_fast.log
