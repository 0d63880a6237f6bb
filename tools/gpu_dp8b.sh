#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n${N}_$name.log 2> gpurun_out/bench_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_$name.log") if l.startswith("{")][-1]); print("N=$N $name", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), round(d["value"]))
except Exception as e:
    print("N=$N $name ERR", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-1200:])
PY
}
run two
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n1_b.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n1_b.log") if l.startswith("{")][-1]); print("N=1", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3))
PY
KR_MULTICAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 tools/dp_check.py 2>&1 | tail -8
