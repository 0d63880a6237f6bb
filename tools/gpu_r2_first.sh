#!/bin/bash
# Round 2, first GPU call: real assertions for the N1/N2/N3 device tests, micro-benchmarks, KR_ATTN_FAST suite, feature ncu.
mkdir -p gpurun_out
python -m pytest tests/test_zz_features_gpu.py tests/test_zz_lengths_gpu.py tests/test_zz_metrics_gpu.py tests/test_zz_inference_gpu.py -m gpu -q --runxfail --tb=long -rfE \
  > gpurun_out/pytest_zz.log 2>&1
tail -15 gpurun_out/pytest_zz.log
timeout 120 python tools/features_bench.py > gpurun_out/features_bench.log 2>&1; cat gpurun_out/features_bench.log
KR_MELSTFT_R4=1 timeout 120 python tools/features_bench.py > gpurun_out/features_bench_r4.log 2>&1; head -1 gpurun_out/features_bench_r4.log
timeout 200 python tools/decode_bench.py 1 64 400 > gpurun_out/decode_bench.log 2>&1; cat gpurun_out/decode_bench.log
KR_DECODE_GEMV=1 timeout 200 python tools/decode_bench.py 1 64 400 > gpurun_out/decode_bench_gemv.log 2>&1; cat gpurun_out/decode_bench_gemv.log
KR_ATTN_FAST=1 python -m pytest tests -m gpu -q -x --deselect tests/test_zz_features_gpu.py --deselect tests/test_zz_inference_gpu.py \
  --deselect tests/test_zz_metrics_gpu.py --deselect tests/test_zz_lengths_gpu.py > gpurun_out/pytest_attn_fast.log 2>&1; tail -3 gpurun_out/pytest_attn_fast.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/features_launches.csv \
  python -c "
import torch, sys
sys.path.insert(0, '.')
from kokoro_ruslan_b200.features import FeaturePipeline, resample
wav = torch.randn(8, 256 * 799 + 100, device='cuda') * 0.1
lens = torch.full((8,), wav.shape[1], dtype=torch.int64, device='cuda')
pipe = FeaturePipeline()
for _ in range(2):
    out = pipe(wav, lens)
    y = resample(wav, 22050, 20506, lengths=lens)
torch.cuda.synchronize()
" > gpurun_out/ncu_features.log 2>&1; tail -2 gpurun_out/ncu_features.log
for k in mel_stft_kernel pitch_frames_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/feat_$k -f \
    python tools/features_bench.py > gpurun_out/ncu_$k.log 2>&1; tail -1 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
