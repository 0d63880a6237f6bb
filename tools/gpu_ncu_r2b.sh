#!/bin/bash
mkdir -p gpurun_out
cap() {  # name regex skip command...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_$name -f "$@" > gpurun_out/ncu_r02_$name.log 2>&1
  tail -1 gpurun_out/ncu_r02_$name.log
}
# one_step.py runs 2 eager steps; a step launches 18 attn_fwd (6 encoder + decoder: causal self / cross alternating) and 18 attn_bwd
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_ --csv --log-file gpurun_out/attn_launches.csv python tools/one_step.py 2 > /dev/null 2>&1
grep -c attn gpurun_out/attn_launches.csv
cap attnfwd_a attn_fwd_kernel 28 python tools/one_step.py 2
cap attnfwd_b attn_fwd_kernel 29 python tools/one_step.py 2
cap attnbwd_a attn_bwd_kernel 20 python tools/one_step.py 2
cap melstft mel_stft_kernel 1 python tools/features_bench.py
timeout 120 python tools/features_bench.py > gpurun_out/features_bench.log 2>&1; cat gpurun_out/features_bench.log
timeout 300 python -m pytest tests/test_melstft_gpu.py tests/test_features_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
ls -la gpurun_out/r02_attn*.ncu-rep gpurun_out/r02_melstft.ncu-rep
