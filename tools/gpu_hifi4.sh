#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_gemm.log 2>&1; tail -15 gpurun_out/pytest_gemm.log
timeout 600 python -m pytest tests/test_hifigan_gpu.py tests/test_parity_configs_gpu.py tests/test_inference_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "hifigan or synthesizer or golden or slab or full_size or refolds or state_dict" > gpurun_out/pytest_hifi.log 2>&1; tail -5 gpurun_out/pytest_hifi.log
timeout 300 python tools/hifigan_bench.py > gpurun_out/hifigan_bench.log 2>&1; tail -2 gpurun_out/hifigan_bench.log
