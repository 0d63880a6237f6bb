"""Micro-benchmark of the row-wise HBM kernels on the decoder critical path at the bench shape (graph replays)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200 import ops
from kokoro_ruslan_b200.params import rope_tables


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


def main():
    N, D, H, S = 6400, 512, 8, 800
    f32 = lambda *s: torch.randn(*s, device="cuda")
    bf = lambda *s: torch.randn(*s, device="cuda").to(torch.bfloat16)
    x, dy, dres, dx = f32(N, D), f32(N, D), f32(N, D), f32(N, D)
    dxb = bf(N, D)
    mean, rstd = f32(N), f32(N).abs() + 0.5
    gam, dgam, dbet = f32(D), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    t = bench(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres, dx, dxb, dgam, dbet))
    print(f"layernorm_bwd  {N}x{D}: {t:6.1f} us  {(4 * N * D * 4 + N * D * 2) / t / 1e3:7.0f} GB/s")
    state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
    rs = torch.ones(8, device="cuda")
    spec = ops.make_drop_spec(state, 3, 0.2, row_scale=rs, rows_per_sample=S)
    t = bench(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam, dres, dx, dxb, dgam, dbet, drop_bf16=spec))
    print(f"layernorm_bwd + dropout on the bf16 copy: {t:6.1f} us")
    cos, sin = rope_tables(4000, 64)
    cos, sin = cos.cuda(), sin.cuda()
    raw, draw = bf(N, 3 * D), bf(N, 3 * D)
    dq, dkv = f32(N, D), bf(N, 2 * D)
    g = [f32(64) for _ in range(3)]
    dg = [torch.zeros(64, device="cuda") for _ in range(3)]
    parts = lambda t_: [t_[:, :D], t_[:, D:2 * D], t_[:, 2 * D:]]
    t = bench(lambda: ops.qkv_prep_bwd(parts(raw), [dq, dkv[:, :D], dkv[:, D:]], parts(draw), g, dg, 0b011, cos, sin, N, S, H))
    byts = N * 3 * D * 2 * 2 + N * D * 4 + N * 2 * D * 2
    print(f"qkv_prep_bwd   3 parts (self-attn): {t:6.1f} us  {byts / t / 1e3:7.0f} GB/s")
    nrm = bf(N, 3 * D)
    t = bench(lambda: ops.qkv_prep_fwd(parts(raw), parts(nrm), g, 0b011, cos, sin, N, S, H))
    print(f"qkv_prep_fwd   3 parts: {t:6.1f} us  {N * 3 * D * 4 / t / 1e3:7.0f} GB/s")
    FF = 1536
    hff, u, du, dh = bf(N, 2 * FF), bf(N, FF), bf(N, FF), bf(N, 2 * FF)
    dspec = ops.make_drop_spec(state, 7, 0.2)
    for nm, sp in (("no dropout", None), ("dropout 0.2", dspec)):
        t = bench(lambda: ops.glu_fwd(hff, u, drop=sp))
        print(f"glu_fwd ({nm}): {t:6.1f} us  {(N * 3 * FF * 2) / t / 1e3:7.0f} GB/s")
        t = bench(lambda: ops.glu_bwd(du, hff, dh, drop=sp))
        print(f"glu_bwd ({nm}): {t:6.1f} us  {(N * 5 * FF * 2) / t / 1e3:7.0f} GB/s")
    hb, mean2, rstd2 = bf(N, D), f32(N), f32(N)
    t = bench(lambda: ops.layernorm_fwd(x, gam, dbet, hb, None, mean2, rstd2))
    print(f"layernorm_fwd: {t:6.1f} us  {(N * D * 6) / t / 1e3:7.0f} GB/s")
    y, out = f32(N, D), f32(N, D)
    t = bench(lambda: ops.rmsnorm_resid_fwd(y, gam, x, out))
    print(f"rmsnorm_resid_fwd: {t:6.1f} us  {(N * D * 12) / t / 1e3:7.0f} GB/s")
    cs = torch.zeros(2 * FF, device="cuda")
    t = bench(lambda: ops.colsum_bf16(dh, cs))
    print(f"colsum_bf16 [6400 x 3072]: {t:6.1f} us  {(N * 2 * FF * 2) / t / 1e3:7.0f} GB/s")
    dyb = bf(N, D)
    t = bench(lambda: ops.rmsnorm_resid_bwd(dy, y, gam, dyb, dgam))
    print(f"rmsnorm_resid_bwd: {t:6.1f} us  {(2 * N * D * 4 + N * D * 2) / t / 1e3:7.0f} GB/s")


if __name__ == "__main__":
    main()
