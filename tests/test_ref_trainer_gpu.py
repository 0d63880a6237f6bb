"""Row (b) of SURVEY.md section 8, end to end: the UNMODIFIED reference KokoroTrainer (baseline/_ref) constructs and trains
the B200 KokoroModel through its own code — _setup_model (with the INTEGRATION.md section 1 import swap), _setup_optimizer
(name-based 10 groups), _setup_ema (copy.deepcopy), _setup_weight_norm_constraints (named_modules lookup), train_epoch
(forward, reference losses, backward through the autograd bridge, pre-clip, explosion detector, clip_grad_norm_,
torch.optim.AdamW on the flat-buffer views, EMA on state_dict views, weight-norm projection) — and lands where the same
trainer lands with the reference's own fp32 CPU model."""
import functools
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_trainer as harness  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600),
              pytest.mark.skipif(not harness.reference_available(), reason="baseline/_ref is not installed")]

OVER = dict(n_mels=80, hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, encoder_ff_dim=256,
            decoder_ff_dim=256, variance_filter_size=64, max_decoder_seq_len=1200, gradient_accumulation_steps=1,
            learning_rate=1e-3, ema_decay=0.9, encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0,
            variance_dropout=0.0, use_stochastic_depth=False, num_epochs=1, use_fused_adamw=False, enable_profiling=False,
            profile_epoch_start=999, use_spec_augment=False, use_torch_compile=False)


def _batches(n=2):
    from oracle import acoustic as oa
    return [oa.synthetic_batch(B=3, P=24, T=150, seed=11 + i, ragged=True) for i in range(n)]


def _seeded_sd():
    from oracle import acoustic as oa
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                             variance_filter=64, max_len=1200)
    return oa.seeded_state_dict(ocfg, seed=0)


def test_reference_trainer_trains_the_b200_model():
    from kokoro_ruslan_b200.model import KokoroModel
    sd = _seeded_sd()
    # the reference arm: same trainer code, the reference's own model, fp32 on the CPU
    ref = harness.build_trainer(None, torch.device("cpu"), 59, OVER)
    ref.model.load_state_dict(sd)
    ref.ema_model.load_state_dict(sd)
    want = harness.run_epoch(ref, _batches())
    # the product arm
    tr = harness.build_trainer(functools.partial(KokoroModel, device="cuda:0"), torch.device("cuda:0"), 59, OVER)
    assert isinstance(tr.model, KokoroModel) and isinstance(tr.model, torch.nn.Module)
    assert isinstance(tr.ema_model, KokoroModel) and tr.ema_model is not tr.model and not tr.ema_model.training
    assert len(tr._dec_ff_weights) == 4 and len(tr._enc_ff_weights) == 4          # named_modules() lookup found them
    assert len(tr.optimizer.param_groups) == 10
    tr.model.load_state_dict(sd)
    tr.ema_model.load_state_dict(sd)
    before = {k: v.detach().clone() for k, v in tr.model.state_dict().items()}
    got = harness.run_epoch(tr, [{k: v.clone() for k, v in b.items()} for b in _batches()])
    assert tr.optimizer_steps_completed == 2 and tr.ema_updates == 2
    for name in ("total_loss", "mel_loss", "dur_loss", "stop_loss"):
        a, b = getattr(got, name), getattr(want, name)
        print(f"{name}: b200 {a:.5f} reference {b:.5f}")
        assert abs(a - b) <= 2e-2 * abs(b) + 1e-4, (name, a, b)
    # the weights moved, in the reference's direction, and the EMA copy follows the 0.9 rule on its own storage
    after, ref_after = tr.model.state_dict(), ref.model.state_dict()
    num = den_a = den_b = 0.0
    for k, w0 in before.items():
        if not torch.is_floating_point(w0) or k.endswith((".pe", "_bins")):
            continue
        da = (after[k].detach().cpu() - w0.cpu()).double().flatten()
        db = (ref_after[k].detach() - sd[k]).double().flatten()
        num += float(da @ db)
        den_a += float(da @ da)
        den_b += float(db @ db)
    cos = num / (den_a ** 0.5 * den_b ** 0.5)
    print(f"update cosine vs the reference trainer on its own model: {cos:.4f}, norms {den_a ** 0.5:.4f} / {den_b ** 0.5:.4f}")
    assert den_a > 0 and cos > 0.8 and abs(den_a ** 0.5 / den_b ** 0.5 - 1.0) < 0.1
    ema = tr.ema_model.state_dict()
    k = "decoder.layers.1.ff.linear1.weight"
    assert not torch.equal(ema[k], after[k]) and not torch.equal(ema[k].cpu(), sd[k])


def test_reference_trainer_spec_augment_closure_is_translated():
    """use_spec_augment=True: train_epoch installs a closure over KokoroTrainer._apply_spec_augment (trainer.py:2049-2055);
    the model translates it into kr_spec_augment spans each forward (never silently ignores it)."""
    from kokoro_ruslan_b200.model import KokoroModel
    over = dict(OVER, use_spec_augment=True, spec_augment_start_epoch=0)
    tr = harness.build_trainer(functools.partial(KokoroModel, device="cuda:0"), torch.device("cuda:0"), 59, over)
    tr.model.load_state_dict(_seeded_sd())
    torch.manual_seed(3)
    got = harness.run_epoch(tr, _batches(2))
    assert tr.optimizer_steps_completed == 2
    assert tr.model._memory_augment_fn is not None and tr.model.engine.spec_spans is not None
    assert got.total_loss == got.total_loss                      # finite
    with pytest.raises(NotImplementedError):
        tr.model.set_memory_augment(lambda mem: mem * 0.5)       # not a zero-mask: refuse, do not ignore
        tr.model(_batches(1)[0]["phoneme_indices"], _batches(1)[0]["mel_specs"], _batches(1)[0]["phoneme_durations"],
                 _batches(1)[0]["stop_token_targets"], pitch_targets=_batches(1)[0]["pitches"],
                 energy_targets=_batches(1)[0]["energies"], stress_indices=_batches(1)[0]["stress_indices"])
