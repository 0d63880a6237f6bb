"""BASELINE.json configs as parity-test cases: config 4 (LengthRegulator stress, T ~ 2000, B = 1) and
config 2 (dynamic frame batching: ragged batches of changing shape through the graph-cached TrainStep)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def _small_cfg(max_len):
    from oracle import acoustic as oa
    from kokoro_ruslan_b200.params import ModelConfig
    ocfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=1, ff_dim=128,
                             variance_filter=64, max_len=max_len)
    cfg = ModelConfig(vocab_size=ocfg.vocab_size, mel_dim=ocfg.mel_dim, hidden_dim=ocfg.hidden_dim,
                      n_encoder_layers=1, n_heads=2, encoder_ff_dim=128, n_decoder_layers=1, decoder_ff_dim=128,
                      max_decoder_seq_len=max_len, variance_filter_size=64, n_variance_bins=ocfg.n_bins)
    return ocfg, cfg


def _batch_from_lengths(mel_lens, seed, P=48):
    """Ragged batch whose durations sum exactly to each utterance's mel length."""
    from oracle import acoustic as oa
    g = torch.Generator().manual_seed(seed)
    B, T = len(mel_lens), max(mel_lens)
    ph_len = torch.randint(P // 2, P + 1, (B,), generator=g)
    ph = torch.zeros(B, P, dtype=torch.long)
    stress = torch.zeros(B, P, dtype=torch.long)
    dur = torch.zeros(B, P, dtype=torch.long)
    for b in range(B):
        n = int(ph_len[b])
        ph[b, :n] = torch.randint(1, 59, (n,), generator=g)
        stress[b, :n] = torch.randint(0, 3, (n,), generator=g)
        cuts = torch.sort(torch.randint(0, mel_lens[b] + 1, (n - 1,), generator=g)).values
        edges = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([mel_lens[b]])])
        dur[b, :n] = edges[1:] - edges[:-1]
    mel = torch.randn(B, T, 80, generator=g) * 2 - 5
    pitch, energy = torch.rand(B, T, generator=g), torch.rand(B, T, generator=g)
    stop = torch.zeros(B, T)
    for b, L in enumerate(mel_lens):
        mel[b, L:] = 0
        pitch[b, L:] = 0
        energy[b, L:] = 0
        stop[b, :L] = oa.build_stop_token_targets(L, tail=6)
    return {"phoneme_indices": ph, "stress_indices": stress, "phoneme_durations": dur, "mel_specs": mel,
            "pitches": pitch, "energies": energy, "stop_token_targets": stop,
            "mel_lengths": torch.tensor(mel_lens), "phoneme_lengths": ph_len}


def test_config4_length_regulator_stress_long_utterance():
    """B = 1, P = 256, T = 2000, durations in [0, 30] incl. zeros: index tensor bit-exact, the step runs with the
    long-sequence stabiliser (loss scale 0.70, clip 0.418) and matches the oracle losses."""
    from oracle import acoustic as oa
    from kokoro_ruslan_b200 import ops
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep
    g = torch.Generator().manual_seed(4)
    P, T = 256, 2000
    cuts = torch.sort(torch.randint(0, T + 1, (P - 1,), generator=g)).values
    edges = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([T])])
    dur = (edges[1:] - edges[:-1]).clamp(max=30).unsqueeze(0)
    dur[0, -1] += T - int(dur.sum())                          # keep the sum at exactly T
    want, L = oa.length_regulate_index(dur)
    idx = torch.empty(1, T, dtype=torch.int32, device="cuda")
    lens = torch.empty(1, dtype=torch.int32, device="cuda")
    ops.lr_index(dur.cuda(), idx, lens)
    assert torch.equal(idx.cpu().long(), want) and int(lens) == T == int(L)

    ocfg, cfg = _small_cfg(2100)
    batch = oa.synthetic_batch(B=1, P=P, T=T, seed=5)
    batch["phoneme_durations"] = dur.clone()
    sd = oa.seeded_state_dict(ocfg, seed=0)
    ts = TrainStep(cfg, sched_cfg=ScheduleConfig(total_steps=10), device="cuda", use_graphs=False)
    ts.load_state_dict(sd)
    losses = ts.train_step({k: v.pin_memory() for k, v in batch.items()}).cpu()
    outs = oa.forward_training(sd, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                               batch["pitches"], batch["energies"], batch["stress_indices"])
    want_l = oa.training_losses(ocfg, outs, batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                                batch["pitches"], batch["energies"], batch["mel_lengths"], batch["phoneme_lengths"])
    for a, b in zip(losses.tolist(), [float(x) for x in want_l]):
        assert abs(a - b) <= 1e-2 * abs(b) + 1e-4, (losses.tolist(), [float(x) for x in want_l])
    ctrl = ts.opt.read_ctrl()
    assert ctrl["step"] == 1 and abs(ctrl["clip_used"] - 0.5 / (2000 / 1400) ** 0.5) < 1e-5


def test_config2_dynamic_batching_changing_shapes():
    """Batches come from DynamicFrameBatchSampler (max_frames = 8000): every shape is new, some repeat; losses of
    every step match the CPU oracle step run on the same sequence of batches."""
    from oracle import acoustic as oa
    from oracle.train_step import CpuTrainStep
    from kokoro_ruslan_b200.data import DynamicFrameBatchSampler
    from kokoro_ruslan_b200.optim import OptimConfig
    from kokoro_ruslan_b200.train_step import ScheduleConfig, TrainStep

    class DS:
        def __init__(self, lens):
            self.samples = [{"audio_length": v} for v in lens]

        def __len__(self):
            return len(self.samples)

    rng = random.Random(3)
    lens = [rng.randint(60, 420) for _ in range(40)]
    random.seed(11)
    sampler = DynamicFrameBatchSampler(DS(lens), max_frames=1600, min_batch_size=1, max_batch_size=6, shuffle=True)
    batches = list(iter(sampler))[:4]
    batches = batches + [batches[0]]                          # a repeated shape exercises the captured graph
    ocfg, cfg = _small_cfg(1200)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    lr = 5e-4
    ts = TrainStep(cfg, OptimConfig(learning_rate=lr), ScheduleConfig(total_steps=100, use_warmup=False), device="cuda",
                   use_graphs=True)
    ts.load_state_dict(sd)
    ref = CpuTrainStep(ocfg, sd, lr=lr)
    for step, idxs in enumerate(batches * 2):
        batch = _batch_from_lengths([lens[i] for i in idxs], seed=100 + sorted(idxs)[0])
        ref.set_lr(ts.sched.lrs()[2])
        want = ref.train_step(batch)
        got = ts.train_step({k: v.pin_memory() for k, v in batch.items()}).cpu().tolist()
        for a, b in zip(got, want):
            assert abs(a - b) <= 2e-2 * abs(b) + 2e-4, (step, got, want)
    assert ts.opt.read_ctrl()["step"] == 2 * len(batches)
