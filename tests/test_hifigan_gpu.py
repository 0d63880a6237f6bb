"""HiFi-GAN generator (tcgen05 implicit-GEMM convs, bf16 activations) vs the fp32 oracle and the golden
audio generated from the live reference.  Tolerance (north star): bf16 within 1e-2, metric
max|a-b| / max|b| on the audio."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-2


def _gen(sd=None, graphs=True):
    from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
    g = HiFiGANGenerator(HiFiGANConfig.get_default_config(), device="cuda", use_graphs=graphs)
    if sd is not None:
        g.load_state_dict(sd, strict=True)
    return g


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("case", ["a", "b"])
def test_matches_reference_golden_audio(case):
    from oracle import hifigan as oh
    fix = np.load(os.path.join(HERE, "golden", "hifigan_small.npz"))
    B, T, seed = [int(v) for v in fix[f"shape_{case}"]]
    sd = oh.seeded_state_dict(oh.HifiConfig(), seed=0)
    gen = _gen(sd)
    mel = oh.synthetic_mel(B, T, seed)
    want = torch.from_numpy(fix[f"audio_{case}"])
    for rep in range(3):                                      # eager warm-up, graph capture, graph replay
        got = gen(mel.cuda()).float().cpu()
        assert got.shape == want.shape
        r = _rel(got, want)
        assert r < TOL, f"case {case} pass {rep}: audio rel err {r:.3e}"
    got_tm = gen(mel.transpose(1, 2).contiguous().cuda()).float().cpu()          # (B,T,80) layout
    assert _rel(got_tm, want) < TOL
    if B == 1:
        got_2d = gen(mel[0].t().contiguous().cuda()).float().cpu()               # (T,80) layout
        assert _rel(got_2d, want) < TOL
    o = oh.generator_forward(sd, oh.HifiConfig(), mel)
    assert _rel(got, o) < TOL


def test_state_dict_keys_match_reference_order():
    from oracle import hifigan as oh
    gen = _gen()
    want = [k for k, _ in oh.state_dict_keys(oh.HifiConfig())]
    assert list(gen.state_dict().keys()) == want
    with pytest.raises(RuntimeError):
        gen.load_state_dict({"conv_pre.weight_g": torch.zeros(1)}, strict=True)


def test_full_size_batch_consistency_and_range():
    """Config 5 size (16 x 800 frames -> 16 x 204800 samples): finite, |audio| <= 1, and item 3 of the
    batch is BIT-identical to running that utterance alone (tiles never mix batch items)."""
    from oracle import hifigan as oh
    sd = oh.seeded_state_dict(oh.HifiConfig(), seed=0)
    gen = _gen(sd)
    mel = oh.synthetic_mel(16, 800, 3).cuda()
    audio = gen(mel).clone()
    assert audio.shape == (16, 1, 204800)
    assert bool(torch.isfinite(audio).all()) and float(audio.abs().max()) <= 1.0
    assert float(audio.std()) > 1e-3
    single = gen(mel[3:4].contiguous()).clone()
    assert torch.equal(single[0], audio[3])
    # locality: changing the last 100 mel frames cannot change audio more than the receptive field away
    mel2 = mel.clone()
    mel2[:, :, 700:] += 1.0
    audio2 = gen(mel2)
    assert torch.equal(audio2[:, :, :(700 - 40) * 256], audio[:, :, :(700 - 40) * 256])
    assert not torch.equal(audio2[:, :, 700 * 256:], audio[:, :, 700 * 256:])


def test_weight_update_refolds():
    from oracle import hifigan as oh
    cfg = oh.HifiConfig()
    gen = _gen(oh.seeded_state_dict(cfg, seed=0))
    mel = oh.synthetic_mel(1, 40, 9)
    a0 = gen(mel.cuda()).clone()
    a0b = gen(mel.cuda()).clone()
    assert torch.equal(a0, a0b)
    sd1 = oh.seeded_state_dict(cfg, seed=1)
    gen.load_state_dict(sd1)
    a1 = gen(mel.cuda()).float().cpu()
    assert _rel(a1, oh.generator_forward(sd1, cfg, mel)) < TOL


def test_slab_path_parity_mid_size():
    """T = 320 frames, B = 1: stages 3 and 4 have >= 296 tiles, so the resident-weight + activation-slab
    conv path is the one that runs; audio still within 1e-2 of the fp32 oracle."""
    from oracle import hifigan as oh
    cfg = oh.HifiConfig()
    sd = oh.seeded_state_dict(cfg, seed=0)
    gen = _gen(sd)
    mel = oh.synthetic_mel(1, 320, 6)
    got = gen(mel.cuda()).float().cpu()
    want = oh.generator_forward(sd, cfg, mel)
    assert _rel(got, want) < TOL


@pytest.mark.parametrize("folded,k,dil,L,final", [(False, 3, 1, 1000, False), (False, 7, 5, 4096 + 77, True),
                                                   (False, 11, 5, 128 * 40 + 3, False), (False, 11, 1, 777, True),
                                                   (True, 3, 3, 2 * 1234, False), (True, 11, 5, 2 * 3000, True),
                                                   (True, 11, 1, 2 * 515, False), (True, 7, 3, 2 * 64, True)])
def test_fused_resblock_step_against_torch(folded, k, dil, L, final):
    """kr_hifi_resblock alone (through ops.hifi_resblock) against torch fp32 on the same bf16-rounded operands: plain
    64-channel steps (k = 11: 44 blocks, one-slot mode) and time-folded 32-channel ones (block-sparse half-block lists),
    ragged lengths, the intermediate (out + out_act) and the final (MRF accumulation, out_act only) epilogue forms."""
    import torch.nn.functional as F
    from kokoro_ruslan_b200 import ops
    from kokoro_ruslan_b200.hifigan import HALO, HiFiGANGenerator as G
    C = 32 if folded else 64
    B = 2
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, L, generator=g)                                          # residual stream (fp32)
    w1 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).to(torch.bfloat16).float()
    w2 = (torch.randn(C, C, k, generator=g) / (C * k) ** 0.5).to(torch.bfloat16).float()
    b1, b2 = torch.randn(C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    xs = torch.randn(B, C, L, generator=g)
    act = F.leaky_relu(x, 0.1).to(torch.bfloat16)
    t = F.leaky_relu(F.conv1d(act.float(), w1, b1, padding=dil * (k - 1) // 2, dilation=dil), 0.1).to(torch.bfloat16).float()
    v = F.conv1d(t, w2, b2, padding=(k - 1) // 2) + x
    if final:
        v = v / 3.0 + xs
    want = v.permute(0, 2, 1)                                                      # [B, L, C]
    x_act = torch.zeros(B, L + 2 * HALO, C, dtype=torch.bfloat16, device="cuda")
    x_act[:, HALO:HALO + L] = act.permute(0, 2, 1).cuda()
    resid = x.permute(0, 2, 1).contiguous().cuda()
    r2 = xs.permute(0, 2, 1).contiguous().cuda()
    out = torch.zeros(B, L, C, device="cuda")
    out_act = torch.zeros(B, L + 2 * HALO, C, dtype=torch.bfloat16, device="cuda")
    wb1, o1, h1 = G._half_blocks(w1.cuda(), dil, folded)
    wb2, o2, h2 = G._half_blocks(w2.cuda(), 1, folded)
    rep = (lambda t_: torch.cat([t_] * 2)) if folded else (lambda t_: t_)
    fl = 2 if folded else 1
    f3 = lambda t_: t_.view(B, t_.shape[1] // fl, 64)                              # noqa: E731  (time-folded view)
    kw = dict(resid=f3(resid), out_act=f3(out_act)[:, HALO // fl:(HALO + L) // fl], act_slope=0.1)
    if final:
        kw.update(resid2=f3(r2), beta=1.0 / 3.0)
    else:
        kw.update(out=f3(out))
    ops.hifi_resblock(f3(x_act), L // fl, HALO // fl, wb1, o1, h1, rep(b1).cuda(), wb2, o2, h2, rep(b2).cuda(), **kw)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    got_act = out_act[:, HALO:HALO + L].float().cpu()
    assert float((got_act - F.leaky_relu(want, 0.1)).abs().max()) / scale < 1e-2
    assert float(out_act[:, :HALO].abs().max()) == 0.0 and float(out_act[:, HALO + L:].abs().max()) == 0.0     # halos untouched
    if not final:
        assert float((out.cpu() - want).abs().max()) / scale < 3e-3
