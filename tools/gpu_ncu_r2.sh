#!/bin/bash
# round 2: ncu --set full captures of the kernels VERDICT r01 asked for (one launch each, second eager step / bench tool)
mkdir -p gpurun_out
cap() {  # name regex skip command...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r02_$name -f "$@" > gpurun_out/ncu_r02_$name.log 2>&1
  tail -1 gpurun_out/ncu_r02_$name.log
}
cap attnfwd_cross "attn_fwd_kernel<false" 14 python tools/one_step.py 2
cap attnfwd_causal "attn_fwd_kernel<true" 8 python tools/one_step.py 2
cap attnbwd "attn_bwd_kernel<false" 14 python tools/one_step.py 2
cap gemm "kr_gemm_kernel" 330 python tools/one_step.py 2
cap adamw "adamw_kernel" 1 python tools/one_step.py 2
cap decattn "dec_attn_kernel" 600 python tools/decode_bench.py 1 64 200
cap decgemv "dec_gemv_kernel" 600 python tools/decode_bench.py 1 64 200
cap hificonv "kr_gemm_kernel" 130 python tools/hifigan_one.py
ls -la gpurun_out/r02_*.ncu-rep
