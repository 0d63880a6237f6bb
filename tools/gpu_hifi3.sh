#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 3" "64 7" "64 11" "128 11"; do
  set -- $cfg
  KR_HIFI_FOLD=$1 KR_HIFI_FOLD_K64=$2 timeout 300 python tools/hifigan_bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fold<=$1 k64<=$2', round(d['ms_per_batch'],3))"
done
KR_HIFI_FOLD=64 KR_HIFI_FOLD_K64=11 timeout 600 python -m pytest tests/test_hifigan_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
