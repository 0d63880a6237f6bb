"""Dropout / stochastic depth on the CUDA path.

torch's Philox stream cannot be reproduced bit for bit, so parity is checked the other way round: the
kernels' masks are pure functions of (seed, step, site, element) and `kr_drop_export_mask` emits them;
the CPU oracle (pinned to the live reference at p = 0) then applies EXACTLY those masks at the reference's
dropout sites (oracle.acoustic `drop` callback) and outputs, losses and all 308 gradients are compared
at the same bf16 tolerances as the deterministic test.  Statistics of the generator are checked
separately.
"""
import math
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _export(state, site, p, rows, cols, ld, byte_lanes=False):
    from kokoro_ruslan_b200 import ops
    out = torch.empty(rows, cols, dtype=torch.uint8, device="cuda")
    ops.drop_export_mask(state, site, p, rows, cols, ld, out, byte_lanes)
    return out


def test_generator_statistics():
    from kokoro_ruslan_b200 import ops
    state = torch.tensor([1234, 7], dtype=torch.int64, device="cuda")
    n = 1 << 22
    for p in (0.1, 0.15, 0.2, 0.5):
        m = _export(state, 3, p, 1, n, n).float()
        keep = ops.drop_keep(p)
        sigma = math.sqrt(keep * (1 - keep) / n)
        assert abs(float(m.mean()) - keep) < 5 * sigma, (p, float(m.mean()), keep)
        assert abs(keep - (1 - p)) < 1e-5
    a = _export(state, 3, 0.5, 1, n, n)
    assert torch.equal(a, _export(state, 3, 0.5, 1, n, n))                       # deterministic
    b = _export(state, 4, 0.5, 1, n, n)                                          # another site
    state2 = torch.tensor([1234, 8], dtype=torch.int64, device="cuda")
    c = _export(state2, 3, 0.5, 1, n, n)                                         # another step
    for other in (b, c):
        agree = float((a == other).float().mean())
        assert abs(agree - 0.5) < 5 * 0.5 / math.sqrt(n), agree
    # neighbouring elements (the two 16-bit lanes of one hash, and consecutive hashes) are uncorrelated
    x = a.float().view(-1) - 0.5
    for lag in (1, 2, 64, 512):
        corr = float((x[:-lag] * x[lag:]).mean() / 0.25)
        assert abs(corr) < 5 / math.sqrt(n), (lag, corr)


@pytest.mark.parametrize("causal", [False, True])
def test_attention_dropout_matches_exported_mask(causal):
    from kokoro_ruslan_b200 import ops
    torch.manual_seed(0)
    B, H, Sq, Sk, p = 2, 2, 200, 200, 0.2
    q, k, v, d_o = (torch.randn(B, S, H, 64, device="cuda").to(torch.bfloat16) for S in (Sq, Sk, Sk, Sq))
    key_mask = torch.zeros(B, Sk, dtype=torch.uint8, device="cuda")
    key_mask[1, 170:] = 1
    state = torch.tensor([5, 3], dtype=torch.int64, device="cuda")
    spec = ops.make_drop_spec(state, 11, p, byte_lanes=True)
    keep_p = 1.0 - ops.drop_thr8(p) / 65536.0
    o = torch.empty_like(q)
    lse = torch.empty(B, H, Sq, device="cuda")
    ops.attn_fwd(q, k, v, o, lse, key_mask, causal, 0.125, drop=spec)
    dq = torch.randn(B, Sq, H, 64, device="cuda")      # kr_attn_bwd zeroes it itself
    dk, dv = torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, H, Sq, device="cuda")
    ops.attn_bwd(q, k, v, o, d_o, lse, delta, dq, dk, dv, key_mask, causal, 0.125, drop=spec)
    torch.cuda.synchronize()

    sk_pad = (Sk + 127) // 128 * 128
    keep = _export(state, 11, p, B * H * Sq, Sk, sk_pad, byte_lanes=True).view(B, H, Sq, Sk).float()
    assert abs(float(keep.mean()) - keep_p) < 5e-3
    qf, kf, vf = (t.float().transpose(1, 2).detach().requires_grad_(True) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.triu(torch.full((Sq, Sk), float("-inf"), device="cuda"), diagonal=1)
    s = s.masked_fill(key_mask.bool().view(B, 1, 1, Sk), float("-inf"))
    pr = torch.softmax(s, dim=-1) * keep / keep_p
    want = pr @ vf
    want.backward(d_o.float().transpose(1, 2))

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    assert rel(o.float().transpose(1, 2), want.detach()) < 1e-2
    assert rel(dv.float().transpose(1, 2), vf.grad) < 1.5e-2
    assert rel(dk.float().transpose(1, 2), kf.grad) < 1.5e-2
    assert rel(dq.transpose(1, 2), qf.grad) < 1.5e-2
    # and the mask really is applied (the un-dropped output differs)
    o0 = torch.empty_like(q)
    ops.attn_fwd(q, k, v, o0, lse, key_mask, causal, 0.125)
    assert rel(o0.float(), o.float()) > 0.05


class _MaskOracle:
    """oracle.acoustic `drop` callback backed by the masks exported from the CUDA library."""

    def __init__(self, eng, B, P, Tp):
        self.eng, self.B, self.P, self.Tp = eng, B, P, Tp
        self.state = eng.drop_state
        self.d = eng.dropout
        self.table = eng._path_table.float().cpu()
        self.seen = set()

    def _mask(self, site, p, rows, cols, ld=None, byte_lanes=False):
        from kokoro_ruslan_b200 import ops
        thr = ops.drop_thr8(p) if byte_lanes else ops.drop_thr(p)
        if thr == 0:
            return torch.ones(rows, cols)
        m = _export(self.state, self.eng.drop_sites[site], p, rows, cols, ld or cols, byte_lanes).float().cpu()
        return m / (1.0 - thr / 65536.0)

    def __call__(self, site, t, **info):
        d = self.d
        self.seen.add(site)
        parts = site.split(".")
        if site in ("enc.pe", "dec.pe"):
            return t * self._mask(site, d.encoder, t.shape[0] * t.shape[1], t.shape[2]).view_as(t)
        if site == "dec.in":
            return t * self._mask(site, d.decoder_input, t.shape[0] * t.shape[1], t.shape[2]).view_as(t)
        if parts[0] == "vp":
            L = self.P if parts[1] == "duration" else self.Tp
            geom = self.eng._geom(self.B, L)
            C, n = t.shape[1], t.shape[2]
            full = self._mask(site, d.variance, geom.R, C)
            rot = geom.row_of_tok.cpu().long().view(self.B, L)[:, info["chunk_start"]:info["chunk_start"] + n]
            return t * full[rot].permute(0, 2, 1)
        p = d.encoder if parts[0] == "enc" else d.decoder
        if parts[-1] == "p":
            Bh, Sq, Sk = t.shape[0] * t.shape[1], t.shape[2], t.shape[3]
            return t * self._mask(site, p, Bh * Sq, Sk, (Sk + 127) // 128 * 128, byte_lanes=True).view_as(t)
        if parts[-1] == "u":
            return t * self._mask(site, p, t.shape[0] * t.shape[1], t.shape[2]).view_as(t)
        assert parts[-1] == "out"
        f = self._mask(site, p, t.shape[0] * t.shape[1], t.shape[2]).view_as(t)
        if parts[2] == "ffn":
            f = f * self._mask(site + "2", p, t.shape[0] * t.shape[1], t.shape[2]).view_as(t)
        path = self.table[self.eng._path_rows[".".join(parts[:3]) + ".path"]]
        return t * f * path.view(-1, 1, 1)


@pytest.mark.parametrize("name", ["tiny", "chunked"])
def test_engine_with_dropout_matches_oracle_under_the_same_masks(name):
    import test_engine_gpu as te
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.engine import DropoutConfig
    ocfg, bk = te._cases()[name]
    batch = oa.synthetic_batch(n_mels=ocfg.mel_dim, vocab=ocfg.vocab_size, **bk)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    eng = te._engine_for(ocfg)
    # stochastic depth 0.5: with B = 2-3 samples and 5-10 branches some (branch, sample) pairs really drop
    eng.set_dropout(DropoutConfig(encoder=0.15, decoder=0.2, decoder_input=0.15, variance=0.1,
                                  stochastic_depth=0.5, seed=77))
    eng.store.load_state_dict(sd)
    outs, ctx, losses, g = te._run_engine(eng, batch)
    eng.zero_grad()
    eng.backward(ctx, g)
    torch.cuda.synchronize()
    tab = eng._path_table.cpu()
    assert ((tab == 0).any() and (tab > 1).any()) or name == "chunked", tab

    cb = _MaskOracle(eng, batch["mel_specs"].shape[0], batch["phoneme_indices"].shape[1], ctx["Tp"])
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    o_outs = oa.forward_training(sdr, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                 batch["pitches"], batch["energies"], batch["stress_indices"], drop=cb)
    o_losses = oa.training_losses(ocfg, o_outs, batch["mel_specs"], batch["phoneme_durations"],
                                  batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                  batch["mel_lengths"], batch["phoneme_lengths"])
    o_losses[0].backward()
    assert {"enc.pe", "dec.in", "dec.pe", "enc.0.attn.p", "enc.0.ffn.u", "dec.0.cross.out", "vp.pitch.1"} <= cb.seen, cb.seen

    errs = {k: te._rel(a.float().cpu(), b.detach()) for k, a, b in
            zip(("mel", "log_dur", "stop", "pitch", "energy"), outs, o_outs)}
    print(name, "dropout output errors vs masked oracle:", errs)
    for k, r in errs.items():
        assert r < (te.TOL_MEL if k == "mel" else te.TOL_OUT_HOT) * 1.5, (k, r)
    got_l = losses.cpu().double()
    want_l = torch.tensor([float(x.detach()) for x in o_losses], dtype=torch.float64)
    assert torch.allclose(got_l, want_l, rtol=1.5e-2, atol=1e-4), (got_l, want_l)
    rows = te._grad_errors(eng, sdr)
    print(name, "worst gradient errors with dropout:", rows[:4])
    te._check_grads(name + " dropout", rows)

    # the deterministic configuration differs (the masks are really applied) ...
    eng.training = False
    outs_eval, _, _, _ = te._run_engine(eng, batch)
    assert te._rel(outs_eval[0].float().cpu(), outs[0].float().cpu()) > 0.02
    # ... and a second training forward draws new masks
    eng.training = True
    outs2, _, _, _ = te._run_engine(eng, batch)
    assert te._rel(outs2[0].float().cpu(), outs[0].float().cpu()) > 0.02


def test_train_step_with_reference_dropout_under_graphs():
    """The full optimizer step with the reference's training dropouts, eager vs CUDA-graph replay: the RNG
    step lives in device memory, so a replayed graph draws fresh masks every step and the two runs agree
    step by step (same seed -> same masks)."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.engine import DropoutConfig
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    cfg = ModelConfig(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                      n_encoder_layers=2, n_heads=2, encoder_ff_dim=256, n_decoder_layers=2, decoder_ff_dim=256,
                      max_decoder_seq_len=1200, variance_filter_size=64, n_variance_bins=ocfg.n_bins)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    hist = []
    for graphs in (False, True):
        ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=100), device="cuda", use_graphs=graphs,
                       dropout=DropoutConfig.reference_training())
        ts.load_state_dict(sd)
        ls = [ts.train_step(batch).cpu().clone() for _ in range(4)]
        hist.append(torch.stack(ls))
        assert int(ts.engine.drop_state[1]) == 4
    assert torch.isfinite(hist[0]).all()
    assert torch.allclose(hist[0], hist[1], rtol=2e-2, atol=1e-3), (hist[0], hist[1])
    assert not torch.allclose(hist[0][0], hist[0][1], rtol=1e-4)      # masks / weights change between steps


def test_masks_bit_exact_against_the_numpy_restatement():
    """The exported CUDA masks equal oracle/dropmask.py bit for bit (integer work: exact), including a padded
    leading dimension (the attention layout) and the stochastic-depth table written by kr_drop_begin."""
    import numpy as np
    from oracle import dropmask
    from kokoro_ruslan_b200 import ops
    state = torch.tensor([987654321, 41], dtype=torch.int64, device="cuda")
    for site, p, rows, cols, ld in [(1, 0.1, 7, 513, 513), (9, 0.2, 33, 200, 256), (200, 0.15, 1, 100001, 100001)]:
        for byte_lanes in (False, True):
            got = _export(state, site, p, rows, cols, ld, byte_lanes).cpu().numpy()
            want = dropmask.keep_mask(987654321, 41, site, p, rows, cols, ld, byte_lanes)
            assert np.array_equal(got, want), (site, p, byte_lanes)
    # kr_drop_begin: advances the step and fills table[s, b] = keep(b) / (1 - p) from site path_site[s]
    sites = torch.tensor([5, 6, 7], dtype=torch.int32, device="cuda")
    probs = torch.tensor([0.0, 0.3, 0.5], dtype=torch.float32, device="cuda")
    table = torch.empty(3, 16, device="cuda")
    ops.drop_begin(state, sites, probs, table, 16)
    assert state.cpu().tolist() == [987654321, 42]
    for s, (site, p) in enumerate([(5, 0.0), (6, 0.3), (7, 0.5)]):
        keep = dropmask.keep_mask(987654321, 42, site, p, 1, 16)[0].astype(np.float32)
        want = keep / (1.0 - dropmask.drop_thr(p) / 65536.0) if p > 0 else np.ones(16, np.float32)
        assert np.allclose(table[s].cpu().numpy(), want, rtol=1e-6), (s, table[s], want)
