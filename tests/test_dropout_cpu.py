"""CPU-side checks of the dropout plumbing: the numpy mask restatement is a sane Bernoulli generator, the host
helpers agree with it, and the oracle's reference-semantics dropout (torch RNG) keeps expectations."""
import math

import numpy as np
import torch


def test_numpy_mask_statistics_and_determinism():
    from oracle import dropmask
    n = 1 << 18
    for p in (0.1, 0.2, 0.5):
        m = dropmask.keep_mask(3, 5, 7, p, 1, n)[0]
        keep = 1.0 - dropmask.drop_thr(p) / 65536.0
        assert abs(m.mean() - keep) < 5 * math.sqrt(keep * (1 - keep) / n)
    a = dropmask.keep_mask(3, 5, 7, 0.5, 4, 1000)
    assert np.array_equal(a, dropmask.keep_mask(3, 5, 7, 0.5, 4, 1000))
    for other in (dropmask.keep_mask(3, 6, 7, 0.5, 4, 1000), dropmask.keep_mask(3, 5, 8, 0.5, 4, 1000),
                  dropmask.keep_mask(4, 5, 7, 0.5, 4, 1000)):
        assert abs((a == other).mean() - 0.5) < 0.05
    # a padded leading dimension only re-indexes the elements
    b = dropmask.keep_mask(3, 5, 7, 0.5, 1, 8 * 1000)[0].reshape(8, 1000)
    assert np.array_equal(dropmask.keep_mask(3, 5, 7, 0.5, 8, 600, 1000), b[:, :600])


def test_host_threshold_helpers_match_the_restatement():
    from oracle import dropmask
    from kokoro_ruslan_b200 import ops
    for p in (0.0, 0.1, 0.15, 0.2, 0.5, 0.999999):
        assert ops.drop_thr(p) == dropmask.drop_thr(p)
        assert abs(ops.drop_keep(p) - (1.0 - dropmask.drop_thr(p) / 65536.0)) < 1e-12
    state = torch.zeros(2, dtype=torch.int64)
    assert ops.make_drop_spec(state, 3, 0.0) is None
    d = ops.make_drop_spec(state, 3, 0.1, 4, 0.2)
    assert (d.site_a, d.site_b) == (3, 4) and abs(d.scale - 1.0 / (ops.drop_keep(0.1) * ops.drop_keep(0.2))) < 1e-6
    d = ops.make_drop_spec(state, 3, 0.0, 4, 0.2)          # a lone second mask moves into slot a
    assert (d.site_a, d.thr_a, d.thr_b) == (4, ops.drop_thr(0.2), 0)


def test_torch_dropout_callback_keeps_expectation_and_drops_paths():
    from oracle import acoustic as oa
    torch.manual_seed(0)
    drop = oa.TorchDropout(p_enc=0.15, p_dec=0.2, p_in=0.15, p_var=0.1, sd_rate=0.5, n_enc=2, n_dec=2)
    t = torch.ones(64, 50, 32)
    for site in ("enc.pe", "dec.in", "enc.0.attn.p", "dec.1.ffn.u", "vp.pitch.0", "enc.0.ffn.out", "dec.1.cross.out"):
        y = drop(site, t)
        assert abs(float(y.mean()) - 1.0) < 0.08, (site, float(y.mean()))
    y = drop("dec.1.self.out", t)                 # layer 1 of 2: drop-path rate 0.5 -> whole samples vanish
    per_sample = y.flatten(1).abs().sum(1)
    assert (per_sample == 0).any() and (per_sample > 0).any()
    y0 = drop("dec.0.self.out", t)                # layer 0: rate 0
    assert (y0.flatten(1).abs().sum(1) > 0).all()


def test_oracle_dropout_sites_reproduce_the_live_reference_under_the_same_torch_rng():
    """tests/golden/acoustic_dropout.npz = the LIVE reference in train() mode with every dropout and stochastic depth
    0.5 on, after torch.manual_seed(seed).  The oracle with the TorchDropout callback consumes the same RNG stream
    (same sites, order, shapes, semantics) and must land on the same outputs: this pins the PLACEMENT of the dropout
    sites that the CUDA path is then checked against (tests/test_dropout_gpu.py, exported masks)."""
    import os
    from oracle import acoustic as oa
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "acoustic_dropout.npz"))
    if str(fix["torch_version"]) != torch.__version__:
        import pytest
        pytest.skip("fixture was generated with another torch RNG implementation")
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                            variance_filter=64, max_len=1200)
    batch = oa.synthetic_batch(n_mels=cfg.mel_dim, vocab=cfg.vocab_size, B=3, P=24, T=150, seed=11, ragged=True)
    sd = oa.seeded_state_dict(cfg, seed=0)
    pe, pd, pi, pv, ps = (float(x) for x in fix["probs"])
    drop = oa.TorchDropout(pe, pd, pi, pv, ps, cfg.n_encoder_layers, cfg.n_decoder_layers)
    torch.manual_seed(int(fix["seed"]))
    outs = oa.forward_training(sd, cfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                               batch["pitches"], batch["energies"], batch["stress_indices"], drop=drop)
    for k, o in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs):
        want = torch.from_numpy(fix[f"out_{k}"])
        assert float((o - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max())), k
    # and the dropouts really fired: the deterministic forward is far away
    det = oa.forward_training(sd, cfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                              batch["pitches"], batch["energies"], batch["stress_indices"])
    assert float((det[0] - torch.from_numpy(fix["out_mel"])).abs().max()) > 0.1
