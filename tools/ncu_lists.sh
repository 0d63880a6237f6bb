#!/bin/bash
# ncu launch lists only (per-launch time + DRAM bytes) of the training step and the HiFi-GAN forward
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/step_launches.csv python tools/one_step.py 2 > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/hifigan_launches.csv python tools/hifigan_one.py > gpurun_out/ncu_hifigan.log 2>&1
tail -1 gpurun_out/ncu_hifigan.log
