"""A few attention forward / backward launches at the decoder shape with dropout (profiled under ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kokoro_ruslan_b200 import ops
B, H, S = 8, 8, 800
mk = lambda: torch.randn(B, S, H, 64, device="cuda").to(torch.bfloat16)
q, k, v, d_o = mk(), mk(), mk(), mk()
o = torch.empty_like(q); lse = torch.empty(B, H, S, device="cuda")
dq = torch.zeros(B, S, H, 64, device="cuda"); dk, dv = torch.empty_like(k), torch.empty_like(v)
delta = torch.empty(B, H, S, device="cuda")
state = torch.tensor([1, 1], dtype=torch.int64, device="cuda")
spec = ops.make_drop_spec(state, 5, 0.2, byte_lanes=True)
for causal in (False, True):
    for _ in range(3):
        ops.attn_fwd(q, k, v, o, lse, None, causal, 0.125, drop=spec)
        ops.attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, None, causal, 0.125, drop=spec)
torch.cuda.synchronize()
print("ok")
