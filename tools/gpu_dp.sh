#!/bin/bash
# usage: gpu_dp.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -x --tb=short -s -p no:cacheprovider > gpurun_out/pytest_dp.log 2>&1; tail -15 gpurun_out/pytest_dp.log
fi
for ov in 1 0; do
  KR_COMM_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n${N}_ov$ov.log 2> gpurun_out/bench_n${N}_ov$ov.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_ov$ov.log") if l.startswith("{")][-1]); print("N=$N overlap=$ov", d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"])
except Exception as e:
    print("N=$N overlap=$ov ERR", e); print(open("gpurun_out/bench_n${N}_ov$ov.err").read()[-1500:])
PY
done
