"""Prints the per-output / per-gradient error table of the CUDA engine vs the fp32 CPU oracle
(diagnostic; run on the GPU box).  usage: python tests/parity_report.py [case ...]  (lives under tests/: it imports the oracle)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))   # this directory

from oracle import acoustic as oa  # noqa: E402
from test_engine_gpu import _cases, _engine_for  # noqa: E402


def report(name):
    ocfg, bk = _cases()[name]
    batch = oa.synthetic_batch(n_mels=ocfg.mel_dim, vocab=ocfg.vocab_size, **bk)
    sd = oa.seeded_state_dict(ocfg, seed=0)
    eng = _engine_for(ocfg)
    eng.store.load_state_dict(sd)
    cb = {k: v.cuda() for k, v in batch.items()}
    outs, ctx = eng.forward(cb["phoneme_indices"], cb["mel_specs"], cb["phoneme_durations"], cb["pitches"],
                            cb["energies"], cb["stress_indices"])
    losses, g = eng.losses(outs, cb["mel_specs"], cb["phoneme_durations"], cb["stop_token_targets"],
                           cb["pitches"], cb["energies"], cb["mel_lengths"], cb["phoneme_lengths"])
    eng.zero_grad()
    eng.backward(ctx, g)
    torch.cuda.synchronize()
    sdr = {k: v.clone().requires_grad_(k not in oa.BUFFER_KEYS) for k, v in sd.items()}
    o_outs = oa.forward_training(sdr, ocfg, batch["phoneme_indices"], batch["mel_specs"], batch["phoneme_durations"],
                                 batch["pitches"], batch["energies"], batch["stress_indices"])
    o_losses = oa.training_losses(ocfg, o_outs, batch["mel_specs"], batch["phoneme_durations"],
                                  batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                  batch["mel_lengths"], batch["phoneme_lengths"])
    o_losses[0].backward()
    print(f"== {name}")
    for key, got, want in zip(("mel", "log_dur", "stop", "pitch", "energy"), outs, o_outs):
        a, b = got.float().cpu(), want.detach()
        print(f"  out {key:8s} max|d|/max|b| {float((a - b).abs().max() / b.abs().max()):.3e}  "
              f"relL2 {float((a - b).norm() / b.norm()):.3e}")
    print("  losses", [round(float(x), 5) for x in losses.cpu()], [round(float(x.detach()), 5) for x in o_losses])
    gsd = eng.store.state_dict(eng.store.grads)
    rows = []
    for n in eng.store.order:
        og = sdr[n].grad
        if og is None:
            continue
        mine = gsd[n].float().cpu()
        e = float((mine - og).norm() / (og.norm() + 1e-20))
        cos = float((mine * og).sum() / (mine.norm() * og.norm() + 1e-20))
        rows.append((e, n, float(og.norm()), cos, float(mine.norm() / (og.norm() + 1e-20))))
    rows.sort(reverse=True)
    for e, n, d, cos, ratio in rows[:int(os.environ.get("TOPN", "25"))]:
        print(f"  grad {e:.3e} cos {cos:.5f} |mine|/|ref| {ratio:.4f} |ref| {d:.3e} {n}")
    import statistics
    print("  median grad err", statistics.median(r[0] for r in rows))


if __name__ == "__main__":
    for c in (sys.argv[1:] or ["tiny", "chunked", "full_width"]):
        report(c)
