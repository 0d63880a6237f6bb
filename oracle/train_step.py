"""ORACLE (test infrastructure, never the product path) — CPU restatement of ONE optimizer step of
the reference trainer on top of oracle.acoustic, in plain fp32 torch.

Order of operations follows the reference (SURVEY.md §8 "Step algorithm" 1-5):
  forward + losses            oracle.acoustic.forward_training / training_losses
  (total * loss_scale).backward()                     training/trainer.py:2284-2294, 3299-3302
  per-tensor spike pre-clip                           training/trainer.py:1332-1407
  clip_grad_norm_(max_norm)                           training/runtime_policies.py:33-79
  AdamW with the 10 parameter groups                  training/trainer.py:446-689
  EMA of the weights                                  training/trainer.py:1491-1517
  FFN weight-norm projection ||W|| <= 95              training/trainer.py:883-912

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it.
Parity pin: the forward/loss/gradient part is pinned to the live reference through
tests/golden/acoustic_*.npz; the optimizer part (group table, pre-clip, explosion detector, clip,
AdamW, EMA, encoder + decoder FFN weight-norm projection) is pinned to the LIVE KokoroTrainer's own
methods through tests/golden/trainer_step.npz (tests/golden/make_golden_trainer_step.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import acoustic as oa

# (group name, lr multiplier, weight decay) at the reference defaults, training/config.py:20-71
GROUPS: List[Tuple[str, float, float]] = [
    ("encoder", 0.65, 0.0), ("encoder_ffn_decay", 0.65, 0.1),
    ("decoder_other_no_decay", 1.0, 0.0), ("decoder_other_decay", 1.0, 0.04),
    ("decoder_attn_decay", 0.15, 0.04), ("decoder_attn_no_decay", 0.15, 0.0),
    ("decoder_ffn_decay", 0.30, 0.35), ("decoder_ffn_no_decay", 0.30, 0.0),
    ("variance_embed", 0.15, 0.0), ("stop_head", 0.10, 0.0),
]


def param_group(name: str) -> int:
    """Index into GROUPS for a parameter name (trainer.py:503-588)."""
    is_bias = name.endswith(".bias")
    normish = any(s in name for s in ("norm.weight", "norm.bias", "layer_norm.weight", "layer_norm.bias"))
    no_decay = is_bias or normish or "duration_adaptor." in name
    encoder = name.startswith(("text_embedding.", "stress_embedding.", "encoder_positional_encoding.",
                               "positional_encoding.", "transformer_encoder_layers.", "encoder_norm."))
    if encoder:
        return 1 if (".ff." in name and not no_decay) else 0
    if name.startswith("stop_token_predictor."):
        return 9
    attn = ".self_attn." in name or ".cross_attn." in name
    if no_decay:
        if "pitch_embedding." in name or "energy_embedding." in name:
            return 8
        return 7 if ".ff." in name else (5 if attn else 2)
    return 6 if ".ff" in name else (4 if attn else 3)


@dataclass
class StepPolicy:
    """Knobs of the optimizer-step boundary at the reference defaults (training/config.py:247-287)."""
    projection_spike_clip_norm: float = 20.0
    attention_spike_clip_norm: float = 4.0
    ffn_spike_clip_norm: float = 3.0
    encoder_ffn_spike_clip_norm: float = 8.0
    stop_head_spike_clip_norm: float = 0.5
    grad_explosion_ema_alpha: float = 0.95
    grad_explosion_abs_floor: float = 1000.0
    grad_explosion_multiplier: float = 3.0
    grad_explosion_warmup_steps: int = 400
    grad_explosion_warmup_floor: float = 8000.0
    grad_explosion_min_ema_steps: int = 100
    emergency_clip: float = 0.3


def spike_clip(name: str, pol: Optional[StepPolicy] = None) -> float:
    """Per-tensor pre-clip ceiling, 0 = none (trainer.py:1340-1392, training/config.py:250-270)."""
    pol = pol or StepPolicy()
    if name.startswith(("mel_projection_in.", "mel_projection_out.")):
        return pol.projection_spike_clip_norm
    if name.startswith("stop_token_predictor."):
        return pol.stop_head_spike_clip_norm
    layered = name.startswith(("decoder.layers.", "transformer_encoder_layers."))
    if layered and name.endswith(".weight") and any(f".{a}.{w}." in name for a in ("self_attn", "cross_attn")
                                                      for w in ("w_q", "w_k", "w_v", "w_o")):
        return pol.attention_spike_clip_norm
    ffn = any(name.endswith(s) for s in (".linear1.weight", ".linear2.weight", ".linear1.bias", ".linear2.bias"))
    if ffn and name.startswith("transformer_encoder_layers."):
        return pol.encoder_ffn_spike_clip_norm
    return pol.ffn_spike_clip_norm if ffn else 0.0


def wn_projected(name: str) -> bool:
    """The 12 decoder AND 12 encoder FFN matrices of the post-step max-norm projection (trainer.py:846-912)."""
    return name.startswith(("decoder.layers.", "transformer_encoder_layers.")) and \
        name.endswith((".ff.linear1.weight", ".ff.linear2.weight"))


class CpuTrainStep:
    """fp32 CPU training step over a flat {name: tensor} state dict."""

    def __init__(self, cfg: oa.AcousticConfig, sd: Dict[str, torch.Tensor], lr: float = 5e-5,
                 max_grad_norm: float = 1.5, ema_decay: float = 0.999, wn_max: float = 95.0, drop=None,
                 policy: Optional[StepPolicy] = None):
        self.cfg = cfg
        self.drop = drop      # oracle.acoustic dropout callback (None = p 0)
        self.sd = {k: (v.clone().requires_grad_(True) if k not in oa.BUFFER_KEYS else v.clone())
                   for k, v in sd.items()}
        self.names = [k for k in self.sd if k not in oa.BUFFER_KEYS]
        groups = [{"params": [], "lr": lr * m, "weight_decay": wd} for _, m, wd in GROUPS]
        for n in self.names:
            groups[param_group(n)]["params"].append(self.sd[n])
        self.opt = torch.optim.AdamW([g for g in groups if g["params"]], betas=(0.9, 0.999), eps=1e-8)
        self.ema = {n: self.sd[n].detach().clone() for n in self.names}
        self.max_grad_norm, self.ema_decay, self.wn_max = max_grad_norm, ema_decay, wn_max
        self.policy = policy or StepPolicy()
        # explosion-detector state (trainer.py:914-925) and the successful-step counter it is keyed on
        self.norm_ema: Optional[float] = None
        self.norm_ema_steps = 0
        self.steps_completed = 0
        self.last: Dict[str, float] = {}

    def set_lr(self, base_lr: float) -> None:
        live = [g for g in GROUPS if any(param_group(n) == GROUPS.index(g) for n in self.names)]
        for pg, (_, m, _) in zip(self.opt.param_groups, live):
            pg["lr"] = base_lr * m

    def fwd_bwd(self, batch: Dict[str, torch.Tensor], loss_scale: float = 1.0):
        for n in self.names:
            self.sd[n].grad = None
        outs = oa.forward_training(self.sd, self.cfg, batch["phoneme_indices"], batch["mel_specs"],
                                   batch["phoneme_durations"], batch["pitches"], batch["energies"],
                                   batch["stress_indices"], drop=self.drop)
        losses = oa.training_losses(self.cfg, outs, batch["mel_specs"], batch["phoneme_durations"],
                                    batch["stop_token_targets"], batch["pitches"], batch["energies"],
                                    batch["mel_lengths"], batch["phoneme_lengths"])
        (losses[0] * loss_scale).backward()
        return outs, losses

    def explosion_threshold(self) -> float:
        """trainer.py:1308-1330: warm-up floor interpolated on the completed optimizer steps; the EMA term joins
        once the detector has seen min_ema_steps norms."""
        pol = self.policy
        floor = pol.grad_explosion_abs_floor
        w = max(0, pol.grad_explosion_warmup_steps)
        if w > 0 and self.steps_completed < w:
            floor = pol.grad_explosion_warmup_floor - (pol.grad_explosion_warmup_floor - floor) * (self.steps_completed / float(w))
        if self.norm_ema_steps < pol.grad_explosion_min_ema_steps or self.norm_ema is None:
            return floor
        return max(floor, self.norm_ema * pol.grad_explosion_multiplier)

    def optimizer_step(self, clip: Optional[float] = None) -> float:
        """The optimizer-step boundary of train_epoch (trainer.py:2345-2470).  Non-finite gradients skip the update
        (trainer.py:2401-2456) — and, unlike the reference, leave the detector's EMA untouched (the reference folds
        the NaN norm into it, after which its threshold silently degrades to the floor for the rest of the run)."""
        pol = self.policy
        clip = self.max_grad_norm if clip is None else clip
        with torch.no_grad():
            for n in self.names:
                g = self.sd[n].grad
                thr = spike_clip(n, pol)
                if g is None or thr <= 0 or not torch.isfinite(g).all():
                    continue
                nrm = float(g.norm(2))
                if nrm > thr:
                    g.mul_(thr / (nrm + 1e-12))
        total = sum(float(self.sd[n].grad.norm(2)) ** 2 for n in self.names if self.sd[n].grad is not None) ** 0.5
        bad = any(self.sd[n].grad is not None and not bool(torch.isfinite(self.sd[n].grad).all()) for n in self.names)
        thr = self.explosion_threshold()
        exploding = (not bad) and total > thr
        if exploding:
            clip = min(clip, pol.emergency_clip)
        self.last = {"total_norm": total, "threshold": thr, "exploding": int(exploding), "clip_used": clip,
                     "skip": int(bad)}
        if bad:
            for n in self.names:
                self.sd[n].grad = None
            return float(total)
        self.norm_ema = total if self.norm_ema is None else \
            pol.grad_explosion_ema_alpha * self.norm_ema + (1 - pol.grad_explosion_ema_alpha) * total
        self.norm_ema_steps += 1
        params = [self.sd[n] for n in self.names]
        total = torch.nn.utils.clip_grad_norm_(params, clip)
        self.opt.step()
        self.steps_completed += 1
        with torch.no_grad():
            for n in self.names:
                self.ema[n].mul_(self.ema_decay).add_(self.sd[n].detach(), alpha=1.0 - self.ema_decay)
            if self.wn_max > 0:
                for n in self.names:
                    if wn_projected(n):
                        nrm = float(self.sd[n].norm(2))
                        if nrm > self.wn_max:
                            self.sd[n].mul_(self.wn_max / nrm)
        return float(total)

    def train_step(self, batch, loss_scale: float = 1.0, clip: Optional[float] = None):
        _, losses = self.fwd_bwd(batch, loss_scale)
        self.optimizer_step(clip)
        return [float(x.detach()) for x in losses]
