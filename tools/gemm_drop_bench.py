"""6400 x 512 x 512 GEMM with the epilogue variants of the attention out-projection (bias + residual, with / without dropout)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200 import ops


def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


M, N, K, S = 6400, 512, 512, 800
a = (torch.randn(M, K, device="cuda") * 0.1).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
bias, resid, out = torch.randn(N, device="cuda"), torch.randn(M, N, device="cuda"), torch.empty(M, N, device="cuda")
state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
rs = torch.ones(8, device="cuda")
print(f"plain fp32 out            : {bench(lambda: ops.gemm(a, w, out)):6.2f} us")
print(f"bias + resid              : {bench(lambda: ops.gemm(a, w, out, bias=bias, resid=resid)):6.2f} us")
for name, spec in (("dropout 0.2", ops.make_drop_spec(state, 3, 0.2)),
                   ("dropout 0.2 + stochastic depth", ops.make_drop_spec(state, 3, 0.2, row_scale=rs, rows_per_sample=S)),
                   ("two masks + stochastic depth", ops.make_drop_spec(state, 3, 0.2, 4, 0.1, row_scale=rs, rows_per_sample=S))):
    print(f"bias + resid + {name:30s}: {bench(lambda: ops.gemm(a, w, out, bias=bias, resid=resid, drop=spec)):6.2f} us")
spec = ops.make_drop_spec(state, 3, 0.2, 4, 0.1, row_scale=rs, rows_per_sample=S)
for bn in (64, 128, 192, 256):
    print(f"two masks + stochastic depth, BLOCK_N {bn:3d}: {bench(lambda: ops.gemm(a, w, out, bias=bias, resid=resid, drop=spec, block_n=bn)):6.2f} us")
