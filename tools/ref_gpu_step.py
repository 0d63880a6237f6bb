"""Like-for-like GPU bar (optional): the UNMODIFIED reference model (installed into baseline/_ref by
`pip install --no-deps --target baseline/_ref /root/reference`) trained for a few steps on the same B200 with plain PyTorch:
bf16 autocast, fused AdamW, the reference's own loss function, gradient checkpointing as the reference configures it.
Same synthetic batch shape as bench.py (B = 8, P = 128, T = 800).  Not part of the product or of bench.py's arms —
it only answers "what does the reference's own GPU path do on this card"."""
import logging
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from bench import B_PER_GPU, N_MELS, P_LEN, T_LEN, synthetic_batch  # noqa: E402
from kokoro.model.model import KokoroModel  # noqa: E402
from kokoro.training.losses import calculate_training_losses  # noqa: E402
from kokoro.utils.lengths import average_by_duration  # noqa: E402

logging.disable(logging.WARNING)
dev = torch.device(os.environ.get("KR_REF_DEV", "cuda"))
if dev.type == "cpu":      # dry run of the script logic only
    B_PER_GPU, P_LEN, T_LEN = 1, 16, 56
torch.manual_seed(0)
model = KokoroModel(vocab_size=59, mel_dim=80, hidden_dim=512, n_encoder_layers=6, n_heads=8, encoder_ff_dim=1536,
                    encoder_dropout=0.15, decoder_dropout=0.20, decoder_input_dropout=0.15, n_decoder_layers=6,
                    decoder_ff_dim=1536, max_decoder_seq_len=4000, variance_filter_size=256, variance_dropout=0.1,
                    n_variance_bins=256, pitch_min=0.0, pitch_max=1.0, energy_min=0.0, energy_max=1.0,
                    use_stochastic_depth=True, stochastic_depth_rate=0.1, qk_norm=True, ffn_output_norm=True).to(dev)
model.train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-5, fused=(dev.type == "cuda"))
batch = {k: v.to(dev) for k, v in synthetic_batch(B_PER_GPU, P_LEN, T_LEN, N_MELS, 59, seed=1).items()}
conf = types.SimpleNamespace(duration_loss_weight=0.35, stop_token_loss_weight=0.01, pitch_loss_weight=1.0,
                             energy_loss_weight=1.0, verbose=False)
crit = dict(criterion_mel=torch.nn.L1Loss(reduction="none"), criterion_duration=torch.nn.HuberLoss(reduction="none", delta=1.0),
            criterion_stop_token=torch.nn.BCEWithLogitsLoss(reduction="none", pos_weight=torch.tensor(17.0, device=dev)),
            criterion_pitch=torch.nn.HuberLoss(reduction="none", delta=0.05),
            criterion_energy=torch.nn.HuberLoss(reduction="none", delta=0.05))
log = logging.getLogger("ref")


def step(autocast: bool):
    opt.zero_grad(set_to_none=True)
    with torch.autocast(dev.type, dtype=torch.bfloat16, enabled=autocast):
        outs = model(batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"], batch["stop_token_targets"],
                     pitch_targets=batch["pitches"], energy_targets=batch["energies"], stress_indices=batch["stress_indices"])
        mel, dur, stop, pitch, energy = outs
        losses = calculate_training_losses(
            device=dev, config=conf, model=model, average_by_duration=average_by_duration, logger=log, predicted_mel=mel,
            predicted_log_durations=dur, predicted_stop_logits=stop, mel_specs=batch["mel_specs"],
            phoneme_durations=batch["phoneme_durations"], stop_token_targets=batch["stop_token_targets"],
            mel_lengths=batch["mel_lengths"], phoneme_lengths=batch["phoneme_lengths"], predicted_pitch=pitch,
            predicted_energy=energy, pitch_targets=batch["pitches"], energy_targets=batch["energies"], **crit)
    losses[0].backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.5)
    opt.step()
    return float(losses[0])


for autocast in (True, False):
    for _ in range(3):
        step(autocast)
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    sync()
    t0 = time.perf_counter()
    n = 8 if dev.type == "cuda" else 1
    for _ in range(n):
        loss = step(autocast)
    sync()
    ms = (time.perf_counter() - t0) / n * 1e3
    import json
    print(json.dumps({"autocast": "bf16" if autocast else "fp32", "ms_per_step": round(ms, 2),
                      "mel_frames_per_s": round(B_PER_GPU * T_LEN / ms * 1e3),
                      "what": "unmodified reference KokoroModel + calculate_training_losses + clip + fused AdamW, plain PyTorch "
                              "eager on this GPU, B=8 P=128 T=800 (no EMA / pre-clip / explosion checks of the trainer)"}))
    print(f"reference PyTorch eager on this GPU, {'bf16 autocast' if autocast else 'fp32'}: {ms:.1f} ms/step = "
          f"{B_PER_GPU * T_LEN / ms * 1e3:,.0f} mel-frames/s (loss {loss:.3f}; fwd+loss+bwd+clip+fused AdamW, no EMA / pre-clip / "
          f"explosion checks of the trainer)")
