#!/bin/bash
# N-GPU confirmation run: the 2-GPU device tests (DP equals the accumulation window, replicas bit-identical), then the bench at N and at 1.
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n${N}_final.log 2> gpurun_out/bench_n${N}_final.err
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n1_final.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/bench_n${N}_final.log", "gpurun_out/bench_n1_final.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["n_gpus"], round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), round(d["value"]), d["config"].get("comm"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/bench_n${N}_final.err
