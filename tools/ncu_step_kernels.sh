#!/bin/bash
# ncu --set full captures of selected kernels inside the second eager training step
mkdir -p gpurun_out
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/step_$1 -f python tools/one_step.py 2 > gpurun_out/ncu_$1.log 2>&1
  tail -2 gpurun_out/ncu_$1.log
}
cap lnbwd ln_bwd_kernel 50
cap prepbwd qkv_prep_bwd_kernel 40
cap colsum colsum_bf16_kernel 80
cap attnfwd "attn_fwd_kernel" 25
cap attnbwd "attn_bwd_kernel" 25
cap gemm "kr_gemm_kernel" 330
ls -la gpurun_out/*.ncu-rep
