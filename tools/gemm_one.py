import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gemm_bench import run
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
bn = int(sys.argv[4]) if len(sys.argv) > 4 else 0
run(M, N, K, out_dtype=torch.bfloat16, bn=bn, iters=2, bias=True)
