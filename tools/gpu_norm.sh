#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/norm_bench.py 2>&1 | grep -E "layernorm_bwd|qkv_prep_bwd|rmsnorm_resid_bwd|glu_bwd"
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -3
bash tools/gpu_ab.sh A=1 A=2
