"""Micro-benchmark of the tcgen05 GEMM at the training-step shapes (back-to-back launches, CUDA events)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kokoro_ruslan_b200 import ops  # noqa: E402
from kokoro_ruslan_b200._lib import check, lib  # noqa: E402


def run(M, N, K, a_mn=False, b_mn=False, out_dtype=torch.float32, splits=1, bn=0, iters=50, bias=False, resid=False):
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda").to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda").to(torch.bfloat16)
    C = torch.zeros(M, N, device="cuda", dtype=out_dtype)
    a = ops.GemmArgs()
    a.A, a.B, a.M, a.N, a.K, a.batch = A.data_ptr(), B.data_ptr(), M, N, K, 1
    a.lda, a.ldb = A.stride(0), B.stride(0)
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.alpha, a.beta = 1.0, 1.0
    bt = torch.randn(N, device="cuda") if bias else None
    rt = torch.randn(M, N, device="cuda") if resid else None
    if bias:
        a.bias = bt.data_ptr()
    if resid:
        a.resid, a.ldr = rt.data_ptr(), N
    a.C, a.ldc = C.data_ptr(), N
    a.c_mode = 2 if splits > 1 else (0 if out_dtype == torch.bfloat16 else 1)
    a.splits, a.force_block_n = splits, bn
    for _ in range(3):
        check(lib().kr_gemm_ex(ctypes.byref(a), ops._stream()))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()          # host launch cost (tensor-map encode, ctypes) must not be timed
    with torch.cuda.graph(g):
        for _ in range(iters):
            check(lib().kr_gemm_ex(ctypes.byref(a), ops._stream()))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
    byts = 2 * (M * K + N * K) + C.numel() * C.element_size() * (2 if splits > 1 else 1)
    print(f"M{M:5d} N{N:5d} K{K:5d} a_mn={int(a_mn)} b_mn={int(b_mn)} out={'bf16' if out_dtype==torch.bfloat16 else 'f32 '} "
          f"splits={splits:2d} bn={bn:3d} bias={int(bias)} resid={int(resid)}: {us:7.2f} us  {tf:7.1f} TF/s  {byts/us/1e3:7.1f} GB/s")
    return us


if __name__ == "__main__":
    for bn in (0, 64, 128, 192, 256):
        run(6400, 512, 512, bn=bn)
    for bn in (0, 128, 256):
        run(6400, 512, 512, bn=bn, out_dtype=torch.bfloat16)
    for bn in (0, 128, 192, 256):
        run(6400, 1536, 512, bn=bn, out_dtype=torch.bfloat16)
    for bn in (0, 128, 192, 256):
        run(6400, 3072, 512, bn=bn, out_dtype=torch.bfloat16, bias=True)
    for bn in (0, 128, 256):
        run(6400, 512, 1536, bn=bn, bias=True)
    run(6400, 512, 512, resid=True, bias=True)
    run(6400, 1536, 512, b_mn=True, out_dtype=torch.bfloat16)
    run(6400, 512, 3072, b_mn=True)
    for sp in (4, 9, 18, 36):
        run(512, 512, 6400, a_mn=True, b_mn=True, splits=sp)
    for sp in (2, 4, 8):
        run(3072, 512, 6400, a_mn=True, b_mn=True, splits=sp)
    run(1024, 512, 512)
    run(1024, 1536, 512, out_dtype=torch.bfloat16)
    run(8192, 8192, 8192, out_dtype=torch.bfloat16, iters=10)
    run(8192, 8192, 8192, out_dtype=torch.bfloat16, iters=10, bn=128)
