"""Conv-mode micro-benchmark at the HiFi-GAN stage-3 shape (16 x 102400 rows, C = 64)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200 import ops

def bench(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

B, L, C, halo = 16, 102400, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 32
x = torch.randn(B, L + 2 * halo, C, device="cuda").to(torch.bfloat16)
out_act = torch.zeros(B, L + 2 * halo, C, device="cuda", dtype=torch.bfloat16)
out_raw = torch.zeros(B, L + 2 * halo, C, device="cuda", dtype=torch.float32)
res = torch.randn(B, L + 2 * halo, C, device="cuda", dtype=torch.float32)
bias = torch.randn(C, device="cuda")
inner = lambda t: t[:, halo:halo + L]
for k in (3, 7, 11):
    w = (torch.randn(C, k * C, device="cuda") / (k * C) ** 0.5).to(torch.bfloat16)
    kw = dict(rows=L, row0=halo - (k - 1) // 2, taps=k, dil=1, bias=bias)
    t_none = bench(lambda: ops.conv1d_cl(x, w, **kw))
    t_act = bench(lambda: ops.conv1d_cl(x, w, out_act=inner(out_act), **kw))
    t_act_ns = bench(lambda: ops.conv1d_cl(x, w, out_act=inner(out_act), no_slab=True, **kw))
    t_full = bench(lambda: ops.conv1d_cl(x, w, out=inner(out_raw), out_act=inner(out_act), resid=inner(res), **kw))
    gf = 2.0 * B * L * C * k * C / 1e9
    print(f"C={C} k={k}: no-store {t_none:7.1f} us | c1 (act only) {t_act:7.1f} us ({gf/t_act/1e3:6.1f} TF/s) | c1 no-slab {t_act_ns:7.1f} | c2 (resid+raw+act) {t_full:7.1f} us")
