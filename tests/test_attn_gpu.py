"""tcgen05 flash attention fwd/bwd vs a plain torch fp32 softmax-attention (oracle S2 core)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, key_mask, causal, scale):
    # q,k,v: [B,S,H,64] fp32 leaf tensors
    qh, kh, vh = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)) * scale
    if causal:
        S = s.shape[-1]
        s = s + torch.triu(torch.full((s.shape[-2], S), float("-inf"), device=s.device), 1)
    if key_mask is not None:
        s = s.masked_fill(key_mask.bool()[:, None, None, :], float("-inf"))
    o = torch.softmax(s, -1) @ vh
    return o.permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,H,Sq,Sk,causal,masked", [
    (2, 2, 128, 128, False, False),
    (2, 8, 800, 800, True, False),
    (2, 8, 800, 800, False, True),
    (3, 2, 150, 150, True, False),
    (3, 2, 37, 37, False, True),
    (1, 8, 128, 320, False, True),
    (1, 4, 2000, 2000, True, False),
])
def test_attention_fwd_bwd(B, H, Sq, Sk, causal, masked):
    from kokoro_ruslan_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + Sq)
    mk = lambda S: torch.randn(B, S, H, 64, generator=g).to(torch.bfloat16).cuda()
    q, k, v, d_o = mk(Sq), mk(Sk), mk(Sk), mk(Sq)
    key_mask = None
    if masked:
        key_mask = torch.zeros(B, Sk, dtype=torch.uint8)
        for b in range(B):
            key_mask[b, Sk - 1 - 7 * b - (Sk // 5):] = 1      # padded tail
            key_mask[b, 3 + b] = 1                            # an interior masked key (id-0 token)
        key_mask = key_mask.cuda()
    scale = 1.0 / math.sqrt(64)
    o = torch.empty_like(q)
    lse = torch.empty(B, H, Sq, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, key_mask, causal, scale)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qf, kf, vf, key_mask, causal, scale)
    err = (o.float() - ref).abs().max().item()
    assert err < 3e-2, f"fwd err {err}"
    ref.backward(d_o.float())
    dq = torch.randn(B, Sq, H, 64, device="cuda")      # kr_attn_bwd zeroes it itself
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, H, Sq, device="cuda")
    ops.attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, key_mask, causal, scale)
    torch.cuda.synchronize()
    for name, got, want in (("dq", dq, qf.grad), ("dk", dk.float(), kf.grad), ("dv", dv.float(), vf.grad)):
        rel = (got - want).abs().max().item() / (want.abs().max().item() + 1e-9)
        assert rel < 3e-2, f"{name} rel err {rel}"


def test_attention_strided_views():
    """q/k/v as column slices of one fused [B*S, 3*H*64] projection output."""
    from kokoro_ruslan_b200 import ops
    B, S, H = 2, 256, 4
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, S, 3 * H * 64, generator=g).to(torch.bfloat16).cuda()
    q = qkv[:, :, :H * 64].view(B, S, H, 64)
    k = qkv[:, :, H * 64:2 * H * 64].view(B, S, H, 64)
    v = qkv[:, :, 2 * H * 64:].view(B, S, H, 64)
    o = torch.empty(B, S, H, 64, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, S, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, None, True, 0.125)
    ref = _ref(q.float(), k.float(), v.float(), None, True, 0.125)
    assert (o.float() - ref).abs().max().item() < 3e-2
