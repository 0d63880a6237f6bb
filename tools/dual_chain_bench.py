"""Would two half-batch forward chains on two streams beat one full-batch chain?  The FFN block chain of the decoder
(LN -> GEMM1 -> GLU -> GEMM2 -> RMSNorm + residual + next LN), graph replays at the decoder shape: one chain over all
6400 rows against two chains over 3200 rows each (row slices of the same buffers) on two captured streams."""
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
side = torch.cuda.Stream()


def chain(r0, r1, layers=6):
    r = slice(r0, r1)
    for _ in range(layers):
        ops.gemm(h[r], w1, hff[r], bias=b1)
        ops.glu_fwd(hff[r], u[r])
        ops.gemm(u[r], w2, y[r], bias=b2)
        ops.rmsnorm_resid_ln_fwd(y[r], gain, x[r], out[r], None, gam, bet, h[r], None, mean[r], rstd[r])


def one():
    chain(0, N)


def two(parts=2):
    cur = torch.cuda.current_stream()
    step = N // parts
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        chain(0, step)
    for i in range(1, parts):
        chain(i * step, (i + 1) * step)
    cur.wait_stream(side)


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
    best = 1e9
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    return best


print(f"6 FFN blocks, one chain over {N} rows: {bench(one):7.1f} us | two chains over {N // 2} rows on two streams: {bench(two):7.1f} us")
