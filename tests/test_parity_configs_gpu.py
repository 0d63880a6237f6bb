"""Parity at the REAL configurations of BASELINE.json, with the measured errors recorded beside the reference's own bf16
noise floor:

  * the bench point (config 1 / 2 synthetic shape): B = 8, P = 128, T = 800, default 6 + 6 layer model at the default
    initialisation — five outputs and six losses against the fp32 oracle; when baseline/_ref is present the LIVE reference
    model runs on the same GPU twice, in fp32 and under bf16 autocast, and every output is gated at
    max(1e-2, 1.25 x the reference's own autocast error) — i.e. at the north-star 1e-2 wherever the reference's bf16 run
    itself stays within 1e-2;
  * config 2 at max_frames = 8000 on the full-width model (ragged dynamic batches through the graph-cached TrainStep);
  * config 5: HiFi-GAN at 2 x 800 frames against the fp32 oracle.

Every measured number is appended to gpurun_out/r02_parity.txt (copied to profiles/ after the run).
"""
import os
import random
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def _record(line: str) -> None:
    print(line)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "r02_parity.txt"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def _live_reference_outputs(ocfg, sd, batch):
    """(fp32 outputs, bf16-autocast outputs) of the unmodified reference KokoroModel on the GPU, or None."""
    if not os.path.isdir(os.path.join(REF, "kokoro")):
        return None
    from oracle.ref_trainer import _import_reference
    _import_reference()                                  # puts baseline/_ref first (and evicts the repo's own kokoro shim)
    import logging
    logging.getLogger("kokoro").setLevel(logging.ERROR)
    from kokoro.model.model import KokoroModel
    m = KokoroModel(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                    n_encoder_layers=ocfg.n_encoder_layers, n_heads=ocfg.n_heads, encoder_ff_dim=ocfg.ff_dim,
                    encoder_dropout=0.0, decoder_dropout=0.0, decoder_input_dropout=0.0,
                    n_decoder_layers=ocfg.n_decoder_layers, decoder_ff_dim=ocfg.ff_dim, max_decoder_seq_len=ocfg.max_len,
                    variance_filter_size=ocfg.variance_filter, variance_dropout=0.0, n_variance_bins=ocfg.n_bins,
                    pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0, use_stochastic_depth=False,
                    qk_norm=True, ffn_output_norm=True, gradient_checkpointing=False)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    cb = {k: v.cuda() for k, v in batch.items()}

    def run():
        return m(cb["phoneme_indices"], cb["mel_specs"], cb["phoneme_durations"], cb["stop_token_targets"],
                 pitch_targets=cb["pitches"], energy_targets=cb["energies"], stress_indices=cb["stress_indices"])
    with torch.no_grad():
        fp32 = [o.float().cpu() for o in run()]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            bf16 = [o.float().cpu() for o in run()]
    del m
    torch.cuda.empty_cache()
    return fp32, bf16


def test_bench_shape_outputs_and_losses_against_oracle_and_reference_noise_floor():
    from kokoro_ruslan_b200.engine import AcousticEngine
    from kokoro_ruslan_b200.params import ModelConfig
    from oracle import acoustic as oa
    ocfg = oa.AcousticConfig(max_len=4000)
    cfg = ModelConfig(max_decoder_seq_len=4000)
    eng = AcousticEngine(cfg, "cuda", with_ema=False)
    eng.store.init_default(seed=0)                     # the reference modules' default initialisation
    sd = {k: v.detach().float().cpu().clone() for k, v in eng.store.state_dict().items()}
    batch = oa.synthetic_batch(B=8, P=128, T=800, seed=21, ragged=True)
    cb = {k: v.cuda() for k, v in batch.items()}
    outs, _ = eng.forward(cb["phoneme_indices"], cb["mel_specs"], cb["phoneme_durations"], cb["pitches"], cb["energies"],
                          cb["stress_indices"])
    losses, _ = eng.losses(outs, cb["mel_specs"], cb["phoneme_durations"], cb["stop_token_targets"], cb["pitches"],
                           cb["energies"], cb["mel_lengths"], cb["phoneme_lengths"])
    got = [o.float().cpu() for o in outs]
    with torch.no_grad():
        o_outs = oa.forward_training(sd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                     batch["pitches"], batch["energies"], batch["stress_indices"])
        o_losses = oa.training_losses(ocfg, o_outs, batch["mel_specs"], batch["phoneme_durations"],
                                      batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                      batch["mel_lengths"], batch["phoneme_lengths"])
    live = _live_reference_outputs(ocfg, sd, batch)
    names = ("mel", "log_dur", "stop", "pitch", "energy")
    _record("# bench shape B=8 P=128 T=800, default 6+6 model, default init: max|a-b|/max|b| per output")
    _record("# output  b200_vs_oracle_fp32  b200_vs_live_ref_fp32  live_ref_bf16_autocast_vs_its_fp32  oracle_vs_live_ref_fp32  gate")
    for i, n in enumerate(names):
        e_or = _rel(got[i], o_outs[i])
        if live is not None:
            e_live, e_ref, e_pin = _rel(got[i], live[0][i]), _rel(live[1][i], live[0][i]), _rel(o_outs[i], live[0][i])
            gate = max(1e-2, 1.25 * e_ref)
            assert e_pin < 1e-3, (n, e_pin)            # the oracle IS the live reference at this size too
        else:
            e_live = e_ref = e_pin = float("nan")
            gate = 1.5e-2
        _record(f"{n:8s} {e_or:.3e} {e_live:.3e} {e_ref:.3e} {e_pin:.3e} {gate:.3e}")
        assert e_or < gate, f"{n}: {e_or:.3e} vs gate {gate:.3e} (reference bf16 autocast: {e_ref:.3e})"
    gl, wl = losses.cpu().tolist(), [float(x) for x in o_losses]
    _record("losses b200   " + " ".join(f"{x:.5f}" for x in gl))
    _record("losses oracle " + " ".join(f"{x:.5f}" for x in wl))
    for a, b in zip(gl, wl):
        assert abs(a - b) <= 1e-2 * abs(b) + 1e-4, (gl, wl)


def test_config2_dynamic_batching_max_frames_8000_full_width_model():
    """Config 2 as BASELINE.json states it: DynamicFrameBatchSampler(max_frames = 8000) on the full-width 6 + 6 model,
    ragged batches whose shapes change every step (one repeats and replays its CUDA graph); per-step losses against the
    CPU oracle step on the same batches."""
    sys.path.insert(0, HERE)
    from test_configs_gpu import _batch_from_lengths
    from oracle import acoustic as oa
    from oracle.train_step import CpuTrainStep
    from kokoro_ruslan_b200.data import DynamicFrameBatchSampler
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.params import ModelConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep

    class DS:
        def __init__(self, lens):
            self.samples = [{"audio_length": v} for v in lens]

        def __len__(self):
            return len(self.samples)

    rng = random.Random(5)
    lens = [rng.randint(300, 1100) for _ in range(48)]
    random.seed(13)
    sampler = DynamicFrameBatchSampler(DS(lens), max_frames=8000, min_batch_size=1, max_batch_size=16, shuffle=True)
    batches = list(iter(sampler))[:2]
    batches = batches + [batches[0]]
    assert max(sum(lens[i] for i in b) for b in batches) > 5000          # the frame budget is really exercised
    ocfg = oa.AcousticConfig(max_len=1200)
    cfg = ModelConfig(max_decoder_seq_len=1200)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    lr = 2e-4
    ts = TrainStep(cfg, OptimConfig(learning_rate=lr), ScheduleConfig(total_steps=100, use_warmup=False), device="cuda",
                   use_graphs=True)
    ts.load_state_dict(sd)
    ref = CpuTrainStep(ocfg, sd, lr=lr)
    for step, idxs in enumerate(batches):
        batch = _batch_from_lengths([lens[i] for i in idxs], seed=200 + sorted(idxs)[0], P=96)
        ref.set_lr(ts.sched.lrs()[2])
        want = ref.train_step(batch)
        got = ts.train_step({k: v.pin_memory() for k, v in batch.items()}).cpu().tolist()
        _record(f"config2 max_frames=8000 step {step}: B={len(idxs)} T={batch['mel_specs'].shape[1]} "
                f"frames={sum(lens[i] for i in idxs)} losses b200 {[round(x, 5) for x in got]} oracle {[round(x, 5) for x in want]}")
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-2 * abs(b) + 2e-4, (step, got, want)
    assert ts.opt.read_ctrl()["step"] == len(batches)


def test_hifigan_2x800_frames_against_oracle():
    """Config 5 content at a size the CPU oracle finishes in seconds: 2 x 800 mel frames -> 2 x 204800 samples."""
    from kokoro_ruslan_b200.hifigan import HiFiGANConfig, HiFiGANGenerator
    from oracle import hifigan as oh
    cfg = oh.HifiConfig()
    sd = oh.seeded_state_dict(cfg, seed=0)
    gen = HiFiGANGenerator(HiFiGANConfig.get_default_config())
    gen.load_state_dict(sd)
    mel = oh.synthetic_mel(2, 800, 7)
    got = gen(mel.cuda()).float().cpu()
    with torch.no_grad():
        want = oh.generator_forward(sd, cfg, mel)
    err = _rel(got, want)
    _record(f"hifigan 2x800 frames: audio max|a-b|/max|b| = {err:.3e} (gate 1e-2)")
    assert got.shape == want.shape == (2, 1, 204800) and err < 1e-2, err
