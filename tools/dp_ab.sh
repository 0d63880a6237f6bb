# A/B of the data-parallel collective on N GPUs (default 2): fused symmetric-memory kernel (multicast / peer) vs NCCL
N=${1:-2}
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 2>gpurun_out/dp_ab_$tag.err | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$tag', 'N=$N', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['config'].get('comm'))" || tail -5 gpurun_out/dp_ab_$tag.err; }
run fused_mc KR_COMM=fused
run fused_peer KR_COMM=fused KR_MULTICAST=0
run nccl KR_COMM=nccl
