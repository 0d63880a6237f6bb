#!/bin/bash
# One profiling pass on the GPU box (1 GPU): launch lists with per-launch time + DRAM bytes of the training step
# and of the HiFi-GAN forward, and `ncu --set full` captures of the top kernels.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/one_step.py 2 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/hifigan_launches.csv python tools/hifigan_one.py > gpurun_out/ncu_hifigan.log 2>&1
tail -2 gpurun_out/ncu_hifigan.log
cap() {  # name regex skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/step_$1 -f python tools/one_step.py 2 > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
cap gemm "kr_gemm_kernel" 330
cap attnfwd "attn_fwd_kernel" 25
cap attnbwd "attn_bwd_kernel" 25
cap adamw "adamw_kernel" 1
ls -la gpurun_out/*.ncu-rep
