#!/bin/bash
N=2
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-hifigan --no-extras > gpurun_out/bench_n${N}_$name.log 2> gpurun_out/bench_n${N}_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n${N}_$name.log") if l.startswith("{")][-1]); print("N=$N $name", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), round(d["value"]), d["config"]["comm"])
except Exception as e:
    print("N=$N $name ERR", e); print(open("gpurun_out/bench_n${N}_$name.err").read()[-1200:])
PY
}
run peer KR_MULTICAST=0
run mc KR_MULTICAST=1
