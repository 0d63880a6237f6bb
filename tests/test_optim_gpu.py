"""Fused clip + AdamW + EMA + projection vs torch.optim.AdamW with the reference's param groups."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch():
    from kokoro_ruslan_b200.optim import FusedAdamW, OptimConfig, group_of, group_hparams, preclip_of
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    cfg = ModelConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1, encoder_ff_dim=128,
                      decoder_ff_dim=128, variance_filter_size=64, max_decoder_seq_len=256)
    store = ParamStore(cfg, torch.device("cuda"))
    store.init_default(seed=1)
    ocfg = OptimConfig(learning_rate=1e-3, ema_decay=0.9, dec_ffn_max_weight_norm=3.0)
    opt = FusedAdamW(store, ocfg)
    # torch reference on CPU copies
    ref = {n: store.ref_view(store.params, n).detach().cpu().clone().requires_grad_(True) for n in store.order}
    hp = group_hparams(ocfg)
    groups = [{"params": [], "lr": ocfg.learning_rate * m, "weight_decay": wd} for m, wd in hp]
    for n in store.order:
        groups[group_of(n)]["params"].append(ref[n])
    topt = torch.optim.AdamW([g for g in groups if g["params"]], betas=ocfg.adam_betas, eps=ocfg.adam_eps)
    ema = {n: ref[n].detach().clone() for n in store.order}
    gen = torch.Generator().manual_seed(7)
    for step in range(3):
        store.grads.zero_()
        for n in store.order:
            gr = torch.randn(ref[n].shape, generator=gen) * (0.05 if step else 2.0)
            ref[n].grad = gr.clone()
            store.ref_view(store.grads, n).copy_(gr.cuda())
        # reference order of operations
        for n in store.order:
            thr = preclip_of(n, ocfg)
            nr = float(ref[n].grad.norm())
            if thr > 0 and nr > thr:
                ref[n].grad.mul_(thr / (nr + 1e-12))
        torch.nn.utils.clip_grad_norm_([ref[n] for n in store.order], ocfg.max_grad_norm)
        topt.step()
        with torch.no_grad():
            for n in store.order:
                ema[n].mul_(ocfg.ema_decay).add_(ref[n].detach(), alpha=1 - ocfg.ema_decay)
            from oracle.train_step import wn_projected            # pinned to the live trainer: decoder AND encoder FFN
            for n in store.order:
                if wn_projected(n):
                    nr = float(ref[n].norm())
                    if nr > ocfg.dec_ffn_max_weight_norm:
                        ref[n].mul_(ocfg.dec_ffn_max_weight_norm / nr)
        opt.step()
    torch.cuda.synchronize()
    ctrl = opt.read_ctrl()
    assert ctrl["step"] == 3 and ctrl["skip"] == 0
    for n in store.order:
        got = store.ref_view(store.params, n).cpu()
        assert torch.allclose(got, ref[n].detach(), rtol=2e-4, atol=2e-6), n
        assert torch.allclose(store.ref_view(store.ema, n).cpu(), ema[n], rtol=2e-4, atol=2e-6), n
        assert torch.allclose(store.ref_view(store.shadow, n).float().cpu(), ref[n].detach(), rtol=1e-2, atol=1e-3), n


def test_nonfinite_gradients_skip_step():
    from kokoro_ruslan_b200.optim import FusedAdamW
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    cfg = ModelConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1, encoder_ff_dim=128,
                      decoder_ff_dim=128, variance_filter_size=64, max_decoder_seq_len=256)
    store = ParamStore(cfg, torch.device("cuda"))
    store.init_default(seed=1)
    before = store.params.clone()
    opt = FusedAdamW(store)
    store.grads.fill_(0.01)
    store.grads[12345] = float("nan")
    opt.step()
    ctrl = opt.read_ctrl()
    assert ctrl["skip"] == 1 and ctrl["step"] == 0 and ctrl["skipped_total"] == 1
    assert torch.equal(store.params, before)


@pytest.mark.parametrize("case", ["plain", "projection", "explosion"])
def test_fused_step_matches_the_live_reference_trainer(case):
    """FusedAdamW (kr_chunk_sqnorm -> kr_step_control -> kr_adamw_step -> kr_wn_project) against
    tests/golden/trainer_step.npz: the LIVE KokoroTrainer's pre-clip, explosion detector, clip, 10-group AdamW, EMA and
    encoder + decoder FFN projection over the same seeded gradients (tests/golden/make_golden_trainer_step.py).  Detector
    decisions exactly, thresholds / norms 1e-5, weights and EMA weights 2e-4 (per-tensor norms and 16 samples each) —
    and every element against the oracle, which reproduces the live trainer bit for bit."""
    import os
    import sys
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden_trainer_step as mk
    from kokoro_ruslan_b200.optim import FusedAdamW, OptimConfig
    from kokoro_ruslan_b200.params import ModelConfig, ParamStore
    from oracle.train_step import CpuTrainStep, StepPolicy
    fx = np.load(os.path.join(here, "golden", "trainer_step.npz"))
    fix = {k.split("/", 1)[1]: fx[k] for k in fx.files if k.startswith(case + "/")}
    spec, c = mk.CASES[case], mk.SMALL
    cfg = ModelConfig(vocab_size=c.vocab_size, mel_dim=c.mel_dim, hidden_dim=c.hidden_dim, n_encoder_layers=c.n_encoder_layers,
                      n_heads=c.n_heads, encoder_ff_dim=c.ff_dim, n_decoder_layers=c.n_decoder_layers, decoder_ff_dim=c.ff_dim,
                      max_decoder_seq_len=c.max_len, variance_filter_size=c.variance_filter, n_variance_bins=c.n_bins)
    store = ParamStore(cfg, torch.device("cuda"))
    sd = mk.scaled_state_dict(case)
    store.load_state_dict(sd)
    opt = FusedAdamW(store, OptimConfig(**spec["over"]))
    pol = StepPolicy(**{k: v for k, v in spec["over"].items() if k.startswith("grad_explosion")})
    ora = CpuTrainStep(c, sd, lr=spec["over"]["learning_rate"], ema_decay=spec["over"]["ema_decay"],
                       wn_max=spec["over"].get("dec_ffn_max_weight_norm", 95.0), policy=pol)
    names = [str(n) for n in fix["names"]]
    assert names == store.order
    shapes = {n: tuple(store.entries[n].shape) for n in names}
    pick = {n: torch.linspace(0, store.entries[n].numel - 1, mk.N_SAMPLES).long() for n in names}
    for step, (gs, special) in enumerate(zip(spec["gscale"], spec["special"])):
        grads = mk.make_grads(names, shapes, case, step, gs)
        if special == "nan":
            grads["decoder.layers.1.ff.linear1.weight"][3, 5] = float("nan")
        store.grads.zero_()
        for n in names:
            store.ref_view(store.grads, n).copy_(grads[n].cuda())
            ora.sd[n].grad = grads[n].clone()
        opt.step()
        ora.optimizer_step()
        ctrl = opt.read_ctrl()
        assert ctrl["skip"] == int(fix["skipped"][step]), (step, ctrl)
        if not ctrl["skip"]:
            assert ctrl["exploding"] == int(fix["exploding"][step]), (step, ctrl)
            assert ctrl["clip_used"] == pytest.approx(float(fix["clip_used"][step]), rel=1e-6)
            assert ctrl["threshold"] == pytest.approx(float(fix["threshold"][step]), rel=1e-5)
            assert ctrl["total_norm"] == pytest.approx(float(fix["total_norm"][step]), rel=1e-5)
        for i, n in enumerate(names):
            for key, buf, osrc in (("w", store.params, ora.sd), ("ema", store.ema, ora.ema)):
                got = store.ref_view(buf, n).detach().cpu()
                flat = got.reshape(-1)
                want_s = torch.from_numpy(fix[key + "_samples"][step][i])
                assert torch.allclose(flat[pick[n]], want_s, rtol=2e-4, atol=2e-6), (step, n, key)
                assert float(flat.double().norm()) == pytest.approx(float(fix[key + "_norms"][step][i]), rel=2e-4, abs=1e-6)
                assert torch.allclose(got, osrc[n].detach(), rtol=2e-4, atol=2e-6), (step, n, key)
    assert opt.read_ctrl()["step"] == len(spec["gscale"]) - sum(1 for x in spec["special"] if x)


def test_batched_conv_dgrad_shadows_equal_the_single_launches():
    """kr_conv_dgrad_shadow_multi (the optimizer step's tail) writes exactly what one kr_conv_dgrad_shadow launch per conv
    writes: Wd[c, j*Co + o] = W[o, 2 - j, c] in bf16."""
    from kokoro_ruslan_b200 import ops
    g = torch.Generator().manual_seed(4)
    Co, Ci = 256, 256
    ws = [torch.randn(Co, 3, Ci, generator=g).cuda() for _ in range(4)]
    single = [torch.zeros(Ci, 3 * Co, dtype=torch.bfloat16, device="cuda") for _ in ws]
    multi = [torch.zeros(Ci, 3 * Co, dtype=torch.bfloat16, device="cuda") for _ in ws]
    for w, wd in zip(ws, single):
        ops.conv_dgrad_shadow(w, wd, Co, Ci)
    ops.conv_dgrad_shadow_multi(list(zip(ws, multi)), Co, Ci)
    torch.cuda.synchronize()
    for w, a, b in zip(ws, single, multi):
        assert torch.equal(a, b)
        want = w.flip(1).permute(2, 1, 0).reshape(Ci, 3 * Co).to(torch.bfloat16)
        assert torch.equal(b, want)
