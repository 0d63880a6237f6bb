#!/bin/bash
# ncu --set full captures of the round's main kernels (final versions); summaries -> gpurun_out/r02_ncu_full_summaries_v2.txt
mkdir -p gpurun_out
cap() {  # name regex skip script
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/r02f_$1 -f python $4 > gpurun_out/ncu_r02f_$1.log 2>&1
  tail -1 gpurun_out/ncu_r02f_$1.log
}
cap rb_k3 hifi_resblock_kernel 0 tools/hifigan_one.py
cap rb_k11 hifi_resblock_kernel 6 tools/hifigan_one.py
cap conv_swap kr_gemm_kernel 14 tools/hifigan_one.py
cap attn_fwd attn_fwd_kernel 2 tools/attn_one.py
cap attn_bwd attn_bwd_kernel 2 tools/attn_one.py
: > gpurun_out/r02_ncu_full_summaries_v2.txt
for n in rb_k3 rb_k11 conv_swap attn_fwd attn_bwd; do
  echo "=== $n ===" >> gpurun_out/r02_ncu_full_summaries_v2.txt
  python tools/ncu_report.py gpurun_out/r02f_$n.ncu-rep >> gpurun_out/r02_ncu_full_summaries_v2.txt 2>&1
done
grep -E "^===|gpu__time_duration|tensor_cycles|dram__bytes_read.sum \[|stall reasons|grid_size|registers_per" gpurun_out/r02_ncu_full_summaries_v2.txt | cut -c1-200
