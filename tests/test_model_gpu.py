"""Drop-in KokoroModel: reference call convention (forward 5-tuple -> torch losses -> loss.backward())
through autograd, parameter names / shapes, SpecAugment on the memory, error behaviour."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(ocfg):
    from kokoro_ruslan_b200.model import KokoroModel
    return KokoroModel(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                       n_encoder_layers=ocfg.n_encoder_layers, n_heads=ocfg.n_heads, encoder_ff_dim=ocfg.ff_dim,
                       encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0,
                       n_decoder_layers=ocfg.n_decoder_layers, decoder_ff_dim=ocfg.ff_dim, max_decoder_seq_len=ocfg.max_len,
                       variance_filter_size=ocfg.variance_filter, variance_dropout=0.0, n_variance_bins=ocfg.n_bins,
                       pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=False,
                       qk_norm=True, ffn_output_norm=True, device="cuda")


def _tiny():
    from oracle import acoustic as oa
    return oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)


@pytest.mark.parametrize("with_spec_augment", [False, True])
def test_reference_call_convention_through_autograd(with_spec_augment):
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.engine import AcousticEngine
    ocfg = _tiny()
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    model = _mk(ocfg)
    model.load_state_dict(sd)
    model.train()
    spans = None
    if with_spec_augment:
        torch.manual_seed(3)
        spans = AcousticEngine.draw_spec_spans(3, batch["mel_specs"].shape[1], ocfg.hidden_dim, 5, 3, 1, 2)
        spans[0, 0, 1] = 4                                    # make sure at least one frame span is non-empty
        spans[1, 1, 1] = 3
        model.set_spec_augment_spans(spans)
    dev = {k: v.cuda() for k, v in batch.items()}
    model.zero_grad(set_to_none=True)
    # exactly how the reference trainer calls the model (trainer.py:3226-3230)
    outs = model(dev["phoneme_indices"], dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
                 pitch_targets=dev["pitches"], energy_targets=dev["energies"], stress_indices=dev["stress_indices"])
    assert len(outs) == 5 and all(o.requires_grad for o in outs)
    losses = oa.training_losses(ocfg, outs, dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
                                dev["pitches"], dev["energies"], dev["mel_lengths"], dev["phoneme_lengths"])
    losses[0].backward()
    torch.cuda.synchronize()
    # oracle
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    o_outs = oa.forward_training(sdr, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                 batch["pitches"], batch["energies"], batch["stress_indices"], spec_spans=spans)
    o_losses = oa.training_losses(ocfg, o_outs, batch["mel_specs"], batch["phoneme_durations"],
                                  batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                  batch["mel_lengths"], batch["phoneme_lengths"])
    o_losses[0].backward()
    for a, b in zip(losses, o_losses):
        assert abs(float(a) - float(b)) <= 1e-2 * abs(float(b)) + 1e-4
    errs = []
    for name, p in model.named_parameters():
        og = sdr[name].grad
        assert p.grad is not None and p.grad.shape == p.shape == sd[name].shape, name
        if og is None or float(og.norm()) < 1e-7:
            continue
        errs.append((float((p.grad.float().cpu() - og).norm() / og.norm()), name))
    errs.sort(reverse=True)
    assert errs[0][0] < 0.15, errs[:5]
    assert sorted(e for e, _ in errs)[len(errs) // 2] < 2.5e-2
    if with_spec_augment:
        # the masked memory changed the outputs relative to the un-augmented forward
        model.set_memory_augment(None)
        plain = model(dev["phoneme_indices"], dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
                      pitch_targets=dev["pitches"], energy_targets=dev["energies"], stress_indices=dev["stress_indices"])
        assert float((plain[0] - outs[0]).abs().max()) > 1e-3


def test_surface_matches_reference_module():
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.model import KokoroModel
    m = KokoroModel(vocab_size=59, encoder_ff_dim=1536, decoder_ff_dim=1536, qk_norm=True, device="cuda")
    names = [n for n, _ in m.named_parameters()]
    assert len(names) == 308 and sum(p.numel() for p in m.parameters()) == 49432276
    assert list(m.state_dict().keys()) == [k for k, _ in oa.state_dict_keys(oa.AcousticConfig())]
    assert len(m.state_dict()) == 311
    info = m.get_model_info()
    assert info["total_parameters"] == 49432276 and info["n_decoder_layers"] == 6
    was = m.training
    x = torch.randint(1, 59, (1, 8)).cuda()
    with pytest.raises(ValueError):
        m(x, torch.zeros(1, 20, 80).cuda())                   # mel without durations / stop targets
    assert m.training == was                                  # forward never flips train/eval
    mel = m.eval()(x)                                         # mel_specs=None -> forward_inference (model.py:813-818)
    assert mel.dim() == 3 and mel.shape[0] == 1 and mel.shape[2] == 80 and mel.shape[1] >= 12
    m.train(was)
    with pytest.raises(RuntimeError):
        m.to("cpu")
    # nn.Module surface the reference trainer relies on (trainer.py:835, 845-881)
    import copy
    mods = dict(m.named_modules())
    assert all(f"decoder.layers.{i}.ff.linear{k}" in mods and f"transformer_encoder_layers.{i}.ff.linear{k}" in mods
               for i in range(6) for k in (1, 2))
    assert mods["decoder.layers.3.ff.linear1"].weight.shape == (3072, 512)
    twin = copy.deepcopy(m)
    assert twin.engine is not m.engine and all(torch.equal(a, b) for a, b in zip(twin.state_dict().values(),
                                                                                 m.state_dict().values()))


def test_explicit_padding_masks_match_oracle_and_reference_golden():
    """forward(..., text_padding_mask=, mel_padding_mask=) (model/model.py:586-589, 648-654): the text mask replaces
    indices == 0 everywhere it is used, the mel mask is the key-padding mask of the decoder self-attention (on top of the
    causal predicate).  Against the oracle and the live-reference fixture; gradients flow."""
    import os
    import numpy as np
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.model import KokoroModel
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    batch = oa.synthetic_batch(B=3, P=24, T=150, seed=11, ragged=True)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    text = batch["phoneme_indices"] == 0
    for b in range(3):
        text[b, int(batch["phoneme_lengths"][b]) - 1] = True
    mel_mask = torch.arange(150).unsqueeze(0) >= batch["mel_lengths"].unsqueeze(1)
    m = KokoroModel(vocab_size=ocfg.vocab_size, hidden_dim=128, n_encoder_layers=2, n_heads=2, encoder_ff_dim=256,
                    n_decoder_layers=2, decoder_ff_dim=256, max_decoder_seq_len=1200, variance_filter_size=64,
                    encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0, variance_dropout=0.0,
                    use_stochastic_depth=False, qk_norm=True, device="cuda")
    m.load_state_dict(sd)
    dev = {k: v.cuda() for k, v in batch.items()}
    outs = m(dev["phoneme_indices"], dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
             pitch_targets=dev["pitches"], energy_targets=dev["energies"], text_padding_mask=text.cuda(),
             mel_padding_mask=mel_mask.cuda(), stress_indices=dev["stress_indices"])
    plain = m(dev["phoneme_indices"], dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
              pitch_targets=dev["pitches"], energy_targets=dev["energies"], stress_indices=dev["stress_indices"])
    with torch.no_grad():
        want = oa.forward_training(sd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                   batch["pitches"], batch["energies"], batch["stress_indices"], text_padding_mask=text,
                                   mel_padding_mask=mel_mask)
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "acoustic_masks.npz"))
    live = ~mel_mask
    for k, got, w in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs, want):
        g = got.detach().float().cpu()
        f = torch.from_numpy(fix[f"out_{k}"])
        if k in ("mel", "stop"):           # frame-level outputs: the real frames (padded ones never reach a loss)
            g, w, f = g[live], w[live], f[live]
        tol = 2e-2 * float(w.abs().max())
        assert float((g - w).abs().max()) < tol and float((g - f).abs().max()) < tol, k
    assert float((plain[0] - outs[0]).abs().max()) > 1e-2          # the masks really changed the forward
    (outs[0].float().pow(2).mean() + outs[1].float().pow(2).mean()).backward()
    gw = dict(m.named_parameters())["decoder.layers.0.self_attn.w_q.weight"].grad
    assert gw is not None and bool(torch.isfinite(gw).all()) and float(gw.abs().max()) > 0
