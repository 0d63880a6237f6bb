#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_step_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -8
bash tools/gpu_ab.sh A=1 A=2
