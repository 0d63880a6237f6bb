#!/bin/bash
# A/B of the row-wise kernels: tools/norm_bench.py lines + the engine tests + the step, per env setting given as arguments
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v"; env $v timeout 200 python tools/norm_bench.py 2>&1 | grep -E "qkv_prep_fwd"
done
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
bash tools/gpu_ab.sh "$@"
