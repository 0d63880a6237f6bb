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
            for n in store.order:
                if n.startswith("decoder.layers.") and (n.endswith("ff.linear1.weight") or n.endswith("ff.linear2.weight")):
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
