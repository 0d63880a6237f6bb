"""HiFi-GAN oracle pinned to the golden audio generated from the live reference module."""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_matches_reference_golden_audio():
    from oracle import hifigan as oh
    fix = np.load(os.path.join(HERE, "golden", "hifigan_small.npz"))
    cfg = oh.HifiConfig()
    sd = oh.seeded_state_dict(cfg, seed=0)
    for case in ("a", "b"):
        B, T, seed = [int(v) for v in fix[f"shape_{case}"]]
        got = oh.generator_forward(sd, cfg, oh.synthetic_mel(B, T, seed))
        want = torch.from_numpy(fix[f"audio_{case}"])
        assert got.shape == want.shape == (B, 1, T * 256)
        assert float((got - want).abs().max()) < 2e-6
        assert float(want.abs().max()) < 0.99          # the fixture is not tanh-saturated


def test_oracle_input_layouts():
    from oracle import hifigan as oh
    cfg = oh.HifiConfig()
    sd = oh.seeded_state_dict(cfg, seed=0)
    mel = oh.synthetic_mel(1, 12, 4)
    a = oh.generator_forward(sd, cfg, mel)
    assert torch.equal(a, oh.generator_forward(sd, cfg, mel.transpose(1, 2)))
    assert torch.equal(a, oh.generator_forward(sd, cfg, mel[0].t()))


def test_state_dict_layout():
    from oracle import hifigan as oh
    keys = oh.state_dict_keys(oh.HifiConfig())
    assert len(keys) == 78 * 3                     # 78 weight-normed convs: bias, g, v
    n = sum(int(np.prod(s)) for k, s in keys if not k.endswith("original0"))
    assert n == 13926017                           # v + biases (g adds one scalar per out/in channel)


def test_vocoder_manager_layout_rules_match_the_reference():
    """mel_to_audio's layout handling (vocoder_manager.py:176-204) on a stub generator: which tensor reaches the
    vocoder and which shape comes back."""
    import torch
    from kokoro_ruslan_b200.hifigan import VocoderManager
    seen = []

    def stub(x):
        seen.append(tuple(x.shape))
        B = x.shape[0]
        T = x.shape[2] if x.shape[1] == 80 else x.shape[1]
        return torch.zeros(B, 1, T * 256)
    vm = VocoderManager("hifigan", device="cpu", vocoder=stub)
    assert tuple(vm.mel_to_audio(torch.zeros(80, 50)).shape) == (12800,) and seen[-1] == (1, 80, 50)
    assert tuple(vm.mel_to_audio(torch.zeros(3, 50, 80)).shape) == (3, 1, 12800) and seen[-1] == (3, 80, 50)
    assert tuple(vm.mel_to_audio(torch.zeros(1, 80, 50)).shape) == (12800,) and seen[-1] == (1, 80, 50)
    assert tuple(vm.mel_to_audio(torch.zeros(1, 50, 80)).shape) == (12800,) and seen[-1] == (1, 50, 80)   # batch of one: untouched
    for bad in ("griffin_lim", "wavernn"):
        try:
            VocoderManager(bad, device="cpu", vocoder=stub)
        except ValueError:
            pass
        else:
            raise AssertionError("expected ValueError")


def test_time_folded_weights_equal_the_dilated_conv():
    """Host logic of the narrow stage (kokoro_ruslan_b200/hifigan.py): a (k, dilation) conv on [L, C] equals the block-sparse
    folded conv with the taps of _folded_taps on the time-folded view [L/2, 2C] — the operand layout the fused ResBlock
    kernel consumes (dilation 1: contiguous taps; dilated: an explicit tap list)."""
    import torch.nn.functional as F
    from kokoro_ruslan_b200.hifigan import HiFiGANGenerator as G
    torch.manual_seed(0)
    C, L = 8, 64
    for k in (3, 7, 11):
        for dil in (1, 3, 5):
            w = torch.randn(C, C, k).to(torch.bfloat16).float()
            x = torch.randn(1, C, L)
            want = F.conv1d(x, w, padding=dil * (k - 1) // 2, dilation=dil)[0]
            taps = G._folded_taps(k, dil)
            assert taps == sorted(taps) and taps[0] == -taps[-1]
            if dil == 1:
                hf = ((k - 1) // 2 + 1) // 2
                assert taps == list(range(-hf, hf + 1))
            wf = G._tap_major_time_folded(w, dil).float().view(2 * C, len(taps), 2 * C)
            xf = x[0].t().contiguous().view(L // 2, 2 * C)
            out = torch.zeros(L // 2, 2 * C)
            for fi, f in enumerate(taps):
                lo, hi = max(0, -f), min(L // 2, L // 2 - f)
                out[lo:hi] += xf[lo + f:hi + f] @ wf[:, fi, :].t()
            assert float((out.view(L, C).t() - want).abs().max()) < 1e-4, (k, dil)


def test_half_block_lists_equal_the_conv():
    """Operand layout of the fused ResBlock kernel (HiFiGANGenerator._half_blocks): summing, over the listed K-half blocks,
    W_i [64 x 32] @ x[row + off_i, 32 kh_i : 32 kh_i + 32] reproduces the conv — plain 64-channel convs and the block-sparse
    time-folded 32-channel ones (zero halves not listed)."""
    import torch.nn.functional as F
    from kokoro_ruslan_b200.hifigan import HiFiGANGenerator as G
    torch.manual_seed(1)
    L = 96
    for folded, C in ((False, 64), (True, 32)):
        for k in (3, 7, 11):
            for dil in (1, 3, 5):
                w = torch.randn(C, C, k).to(torch.bfloat16).float()
                x = torch.randn(1, C, L)
                want = F.conv1d(x, w, padding=dil * (k - 1) // 2, dilation=dil)[0].t()          # [L, C]
                wb, offs, khs = G._half_blocks(w, dil, folded)
                n = len(offs)
                assert wb.shape == (64, n * 32) and len(khs) == n and set(khs) <= {0, 1}
                rows = L // 2 if folded else L
                xv = x[0].t().contiguous().view(rows, 64)
                out = torch.zeros(rows, 64)
                for i in range(n):
                    blk = wb[:, i * 32:(i + 1) * 32].float()
                    lo, hi = max(0, -offs[i]), min(rows, rows - offs[i])
                    out[lo:hi] += xv[lo + offs[i]:hi + offs[i], khs[i] * 32:(khs[i] + 1) * 32] @ blk.t()
                assert float((out.view(L, C) - want).abs().max()) < 1e-4, (folded, k, dil)
                if folded:      # the block-sparse form lists at most 2 blocks per original tap
                    assert n <= 2 * k
    # all nine steps of the folded 32-channel stage fit next to the kernel's 83 KB of slabs (35 blocks of 4 KB)
    for k in (3, 7, 11):
        for dil in (1, 3, 5):
            n1 = len(G._half_blocks(torch.zeros(32, 32, k), dil, True)[1])
            n2 = len(G._half_blocks(torch.zeros(32, 32, k), 1, True)[1])
            assert n1 + n2 <= 35, (k, dil, n1, n2)


def test_oracle_equals_the_installed_reference_on_edge_shapes():
    """Side by side with the installed reference generator (baseline/_ref) on shapes the fixture does not hold: one frame, an
    odd handful of frames, a batch of one and of three, an unbatched (80, T) mel, loud / silent inputs."""
    import logging
    import sys
    import pytest
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import hifigan as oh
    from oracle import ref_trainer
    if not ref_trainer.reference_available():
        pytest.skip("baseline/_ref is not installed")
    ref_trainer._import_reference()
    logging.getLogger("kokoro").setLevel(logging.CRITICAL)
    from kokoro.inference.hifigan_vocoder import HiFiGANConfig, HiFiGANGenerator
    cfg = oh.HifiConfig()
    ref = HiFiGANGenerator(HiFiGANConfig.get_default_config()).eval()
    sd = oh.seeded_state_dict(cfg, seed=0)
    ref.load_state_dict(sd, strict=True)
    cases = {"one frame": oh.synthetic_mel(1, 1, 5), "seven frames, three utterances": oh.synthetic_mel(3, 7, 6),
             "unbatched (T, 80)": oh.synthetic_mel(1, 9, 7)[0].t(), "silence": torch.full((1, 80, 5), -11.5),
             "loud": oh.synthetic_mel(2, 6, 8) + 6.0}
    for label, mel in cases.items():
        with torch.no_grad():
            want = ref(mel)
        got = oh.generator_forward(sd, cfg, mel)
        assert got.shape == want.shape, (label, got.shape, want.shape)
        assert float((got - want).abs().max()) < 2e-6, (label, float((got - want).abs().max()))
    # a 2-D input is ALWAYS read as (time, mel) (hifigan_vocoder.py:115-117): (80, T) is an error on both sides
    wrong = oh.synthetic_mel(1, 9, 7)[0]
    with pytest.raises(RuntimeError):
        ref(wrong)
    with pytest.raises(RuntimeError):
        oh.generator_forward(sd, cfg, wrong)
