"""Marginal cost of rmsnorm_resid_fwd inside the FFN chain it lives in (LN -> GEMM1 -> GLU -> GEMM2 -> RMSNorm + resid),
graph replays at the decoder shape, with / without the kernel and with / without its dropout spec."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200 import ops

N, D, FF, S = 6400, 512, 1536, 800
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.1).to(torch.bfloat16)
f32 = lambda *s: torch.randn(*s, device="cuda")
x, gam, bet = f32(N, D), f32(D), f32(D)
h, hff, u = bf(N, D), bf(N, 2 * FF), bf(N, FF)
w1, b1, w2, b2 = bf(2 * FF, D), f32(2 * FF), bf(D, FF), f32(D)
y, out, gain = f32(N, D), f32(N, D), f32(D)
mean, rstd = f32(N), f32(N)
state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
rs = torch.ones(8, device="cuda")
spec = ops.make_drop_spec(state, 3, 0.2, 4, 0.1, row_scale=rs, rows_per_sample=S)


def chain(rms: bool, drop, layers=6):
    for _ in range(layers):
        ops.layernorm_fwd(x, gam, bet, h, None, mean, rstd)
        ops.gemm(h, w1, hff, bias=b1)
        ops.glu_fwd(hff, u)
        ops.gemm(u, w2, y, bias=b2)
        if rms:
            ops.rmsnorm_resid_fwd(y, gain, x, out, drop=drop)


def bench(fn, iters=5):
    for _ in range(2):
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


a = bench(lambda: chain(False, None))
b = bench(lambda: chain(True, None))
c = bench(lambda: chain(True, spec))
print(f"6 FFN blocks without rmsnorm_resid_fwd {a:7.1f} us | with {b:7.1f} us (+{(b - a) / 6:5.1f} us each) | with dropout spec {c:7.1f} us (+{(c - a) / 6:5.1f} us each)")
d = bench(lambda: [ops.rmsnorm_resid_fwd(y, gain, x, out, drop=spec) for _ in range(6)])
e = bench(lambda: [ops.rmsnorm_resid_fwd(y, gain, x, out) for _ in range(6)])
print(f"rmsnorm_resid_fwd alone: {d / 6:5.1f} us with dropout, {e / 6:5.1f} us without")
