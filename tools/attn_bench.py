"""Micro-benchmark of the flash-attention kernels at the decoder shapes (CUDA-graph replays)."""
import math, os, sys
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


def run(B, H, S, causal):
    mk = lambda: torch.randn(B, S, H, 64, device="cuda").to(torch.bfloat16)
    q, k, v, d_o = mk(), mk(), mk(), mk()
    o = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
    dq = torch.zeros(B, S, H, 64, device="cuda"); dk, dv = torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, H, S, device="cuda")
    f = bench(lambda: ops.attn_fwd(q, k, v, o, lse, None, causal, 0.125))
    b = bench(lambda: ops.attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, None, causal, 0.125))
    import ctypes
    from kokoro_ruslan_b200._lib import lib
    def bwd_nodq():
        lib().kr_attn_bwd(ops._ptr(q), *ops._heads_strides(q), ops._ptr(k), *ops._heads_strides(k), ops._ptr(v),
                          *ops._heads_strides(v), ops._ptr(o), *ops._heads_strides(o), ops._ptr(d_o), *ops._heads_strides(d_o),
                          ops._ptr(lse), ops._ptr(delta), None, ctypes.c_longlong(dq.stride(1)), ctypes.c_longlong(dq.stride(0)),
                          ops._ptr(dk), *ops._heads_strides(dk), ops._ptr(dv), *ops._heads_strides(dv), None,
                          ctypes.c_int(B), ctypes.c_int(H), ctypes.c_int(S), ctypes.c_int(S), ctypes.c_int(int(causal)),
                          ctypes.c_float(0.125), None, ops._stream())
    b2 = bench(bwd_nodq)
    print(f"   bwd without dQ atomics: {b2:7.1f} us")
    state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
    spec = ops.make_drop_spec(state, 5, 0.2, byte_lanes=True)
    fd = bench(lambda: ops.attn_fwd(q, k, v, o, lse, None, causal, 0.125, drop=spec))
    bd = bench(lambda: ops.attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, None, causal, 0.125, drop=spec))
    print(f"   with dropout 0.2: fwd {fd:7.1f} us | bwd {bd:7.1f} us")
    mm = 2.0 * B * H * S * S * 64 * (0.5 if causal else 1.0)
    print(f"B{B} H{H} S{S} causal={int(causal)}: fwd {f:7.1f} us {2*mm/f/1e6:6.1f} TF/s | bwd {b:7.1f} us {5*mm/b/1e6:6.1f} TF/s")


if __name__ == "__main__":
    run(8, 8, 800, False)
    run(8, 8, 800, True)
    run(8, 8, 128, False)
    run(1, 8, 2000, True)
