"""Fused multi-tensor optimizer over the flat parameter store.

Reproduces, with four kernel launches and zero host synchronisations, the reference's per-step
sequence (SURVEY.md §8 "Step algorithm" 4-5): per-tensor spike pre-clip -> total norm ->
explosion detector -> clip_grad_norm_ -> 10-group AdamW -> EMA -> FFN weight-norm projection.
Grouping rules restate ``KokoroTrainer._setup_optimizer`` (reference
src/kokoro/training/trainer.py:446-689) for an UN-compiled model, i.e. the intended 10 groups.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from ._lib import check, lib
from .ops import _ptr, _stream, c_float, c_int
from .params import ParamStore

CHUNK = 4096


@dataclass
class OptimConfig:
    """reference training/config.py:20-71,247-287,339-343"""
    learning_rate: float = 5.0e-5
    weight_decay: float = 0.04
    ffn_weight_decay: float = 0.1
    decoder_ffn_weight_decay: float = 0.35
    encoder_lr_multiplier: float = 0.65
    stop_head_lr_multiplier: float = 0.1
    decoder_ffn_lr_multiplier: float = 0.30
    decoder_attn_lr_multiplier: float = 0.15
    variance_embedding_lr_multiplier: float = 0.15
    adam_eps: float = 1e-8
    adam_betas: Tuple[float, float] = (0.9, 0.999)
    max_grad_norm: float = 1.5
    projection_spike_clip_norm: float = 20.0
    attention_spike_clip_norm: float = 4.0
    ffn_spike_clip_norm: float = 3.0
    encoder_ffn_spike_clip_norm: float = 8.0
    stop_head_spike_clip_norm: float = 0.5
    dec_ffn_max_weight_norm: float = 95.0
    grad_explosion_warmup_steps: int = 400
    grad_explosion_warmup_floor: float = 8000.0
    grad_explosion_abs_floor: float = 1000.0
    grad_explosion_min_ema_steps: int = 100
    grad_explosion_ema_alpha: float = 0.95
    grad_explosion_multiplier: float = 3.0
    emergency_clip: float = 0.3
    ema_decay: float = 0.999


def recommended_ema_decay(n_train: int, batch_size: int, k: float) -> float:
    """EMA decay whose half-life is k epochs (reference utils/ema.py:6-27): exp(-ln 2 / (n_train / batch_size * k)),
    clipped to [0.9, 0.9999]; 0.9999 for degenerate inputs.  The reference trainer calls it with the optimizer steps per
    epoch as n_train and batch_size = 1 (training/trainer.py:808-822) when config.ema_decay is None — the default."""
    if n_train <= 0 or batch_size <= 0:
        return 0.9999
    half_life_steps = (n_train / batch_size) * k
    if half_life_steps <= 0:
        return 0.9999
    return max(0.9, min(math.exp(-math.log(2) / half_life_steps), 0.9999))


GROUP_NAMES = ["encoder", "encoder_ffn_decay", "decoder_other_no_decay", "decoder_other_decay",
               "decoder_attn_decay", "decoder_attn_no_decay", "decoder_ffn_decay", "decoder_ffn_no_decay",
               "variance_embed", "stop_head"]

_ENC_PREFIXES = ("text_embedding.", "stress_embedding.", "encoder_positional_encoding.", "positional_encoding.",
                 "transformer_encoder_layers.", "encoder_norm.")
_NO_DECAY_SUBSTR = ("norm.weight", "norm.bias", "layer_norm.weight", "layer_norm.bias", "duration_adaptor.")


def group_of(name: str) -> int:
    """Index into GROUP_NAMES (trainer.py:503-588)."""
    no_decay = name.endswith(".bias") or any(s in name for s in _NO_DECAY_SUBSTR)
    if any(name.startswith(p) for p in _ENC_PREFIXES):
        return 1 if (".ff." in name and not no_decay) else 0
    if name in ("stop_token_predictor.weight", "stop_token_predictor.bias"):
        return 9
    if no_decay:
        if "pitch_embedding." in name or "energy_embedding." in name:
            return 8
        if ".ff." in name:
            return 7
        if ".self_attn." in name or ".cross_attn." in name:
            return 5
        return 2
    if ".ff." in name or ".ff" in name:
        return 6
    if ".self_attn." in name or ".cross_attn." in name:
        return 4
    return 3


def group_hparams(cfg: OptimConfig) -> List[Tuple[float, float]]:
    """(lr multiplier, weight decay) per group."""
    return [(cfg.encoder_lr_multiplier, 0.0), (cfg.encoder_lr_multiplier, cfg.ffn_weight_decay),
            (1.0, 0.0), (1.0, cfg.weight_decay),
            (cfg.decoder_attn_lr_multiplier, cfg.weight_decay), (cfg.decoder_attn_lr_multiplier, 0.0),
            (cfg.decoder_ffn_lr_multiplier, cfg.decoder_ffn_weight_decay), (cfg.decoder_ffn_lr_multiplier, 0.0),
            (cfg.variance_embedding_lr_multiplier, 0.0), (cfg.stop_head_lr_multiplier, 0.0)]


_ATTN_FRAGS = tuple(f".{a}.{w}.weight" for a in ("self_attn", "cross_attn") for w in ("w_q", "w_k", "w_v", "w_o"))
_FFN_FRAGS = (".linear1.weight", ".linear2.weight", ".linear1.bias", ".linear2.bias")


def preclip_of(name: str, cfg: OptimConfig) -> float:
    """Per-tensor spike pre-clip threshold, 0 = none (trainer.py:1332-1407)."""
    if name.startswith("mel_projection_in.") or name.startswith("mel_projection_out."):
        return cfg.projection_spike_clip_norm
    if name.startswith("stop_token_predictor."):
        return cfg.stop_head_spike_clip_norm
    if (name.startswith("decoder.layers.") or name.startswith("transformer_encoder_layers.")) and \
            any(f in name for f in _ATTN_FRAGS):
        return cfg.attention_spike_clip_norm
    if name.startswith("transformer_encoder_layers.") and any(f in name for f in _FFN_FRAGS):
        return cfg.encoder_ffn_spike_clip_norm
    if any(f in name for f in _FFN_FRAGS):
        return cfg.ffn_spike_clip_norm
    return 0.0


def wnmax_of(name: str, cfg: OptimConfig) -> float:
    """decoder.layers.{i}.ff.linear{1,2}.weight AND transformer_encoder_layers.{i}.ff.linear{1,2}.weight (12 + 12 matrices
    at the default depth) are projected to ||W|| <= dec_ffn_max_weight_norm after every successful step
    (trainer.py:845-881 registers both lists, :900-912 clamps both with the same ceiling)."""
    if (name.startswith("decoder.layers.") or name.startswith("transformer_encoder_layers.")) and \
            (name.endswith(".ff.linear1.weight") or name.endswith(".ff.linear2.weight")):
        return max(0.0, cfg.dec_ffn_max_weight_norm)
    return 0.0


CTRL_FIELDS = ["total_norm", "clip_coef", "skip", "step", "bc1", "bc2_sqrt", "ema_norm", "ema_steps",
               "exploding", "threshold", "nonfinite", "clip_used", "skipped_total"]
_CTRL_INT = {"skip", "step", "ema_steps", "exploding", "nonfinite", "skipped_total"}


class FusedAdamW:
    def __init__(self, store: ParamStore, cfg: Optional[OptimConfig] = None):
        self.store = store
        self.cfg = cfg or OptimConfig()
        dev = store.device
        names = store.order
        self.n_tensors = len(names)
        t_group = [group_of(n) for n in names]
        chunk_tensor, chunk_start, chunk_len, wn_chunks = [], [], [], []
        t_wn = [wnmax_of(n, self.cfg) for n in names]
        for ti, n in enumerate(names):
            e = store.entries[n]
            for s in range(0, e.numel, CHUNK):
                if t_wn[ti] > 0:
                    wn_chunks.append(len(chunk_tensor))
                chunk_tensor.append(ti)
                chunk_start.append(e.offset + s)
                chunk_len.append(min(CHUNK, e.numel - s))
        i32 = dict(dtype=torch.int32, device=dev)
        self.t_group = torch.tensor(t_group, **i32)
        self.chunk_tensor = torch.tensor(chunk_tensor, **i32)
        self.chunk_start = torch.tensor(chunk_start, dtype=torch.int64, device=dev)
        self.chunk_len = torch.tensor(chunk_len, **i32)
        self.wn_chunks = torch.tensor(wn_chunks if wn_chunks else [0], **i32)
        self.n_wn_chunks = len(wn_chunks)
        self.n_chunks = len(chunk_tensor)
        self.t_preclip = torch.tensor([preclip_of(n, self.cfg) for n in names], dtype=torch.float32, device=dev)
        self.t_wnmax = torch.tensor(t_wn, dtype=torch.float32, device=dev)
        first_chunk = [0] * (self.n_tensors + 1)
        for ti in chunk_tensor:
            first_chunk[ti + 1] += 1
        for ti in range(self.n_tensors):
            first_chunk[ti + 1] += first_chunk[ti]
        self.first_chunk = torch.tensor(first_chunk, **i32)
        # per-chunk squared gradient sums: written by kr_chunk_sqnorm, or — data parallel — by the fused
        # all-reduce kernel, in which case the buffer lives in symmetric memory (parallel.SymmetricGradReducer)
        self.sq_chunk = torch.zeros(self.n_chunks, dtype=torch.float32, device=dev)
        self.sq = torch.zeros(self.n_tensors, dtype=torch.float32, device=dev)
        self.wsq = torch.zeros(self.n_tensors, dtype=torch.float32, device=dev)
        self.tscale = torch.ones(self.n_tensors, dtype=torch.float32, device=dev)
        hp = group_hparams(self.cfg)
        self.lr_mult = [m for m, _ in hp]
        self.g_wd = torch.tensor([w for _, w in hp], dtype=torch.float32, device=dev)
        self.g_lr = torch.tensor([self.cfg.learning_rate * m for m in self.lr_mult], dtype=torch.float32, device=dev)
        # pinned ring: the async H2D of step n must not observe the host write of step n+1
        self._lr_ring = torch.empty(64, len(hp), dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None
        self._lr_i = 0
        assert lib().kr_optim_ctrl_size() == 64
        self.ctrl = torch.zeros(16, dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------------------------------
    def set_lrs(self, lrs: List[float]) -> None:
        """Per-group learning rates for the next step (scheduler output)."""
        slot = self._lr_ring[self._lr_i % self._lr_ring.shape[0]]
        self._lr_i += 1
        slot.copy_(torch.tensor(lrs, dtype=torch.float32))
        self.g_lr.copy_(slot, non_blocking=True)

    def set_base_lr(self, base_lr: float) -> None:
        self.set_lrs([base_lr * m for m in self.lr_mult])

    def step(self, clip_norm: Optional[float] = None, clip_override: Optional[torch.Tensor] = None,
             sq_chunk_ready: bool = False) -> None:
        """One optimizer step on store.grads (already all-reduced when data-parallel).  sq_chunk_ready: the
        per-chunk squared sums were already produced by the fused all-reduce kernel."""
        c, s = self.cfg, self.store
        L = lib()
        from .ops import zero_
        zero_(self.wsq)
        if not sq_chunk_ready:
            check(L.kr_chunk_sqnorm(_ptr(s.grads), _ptr(self.chunk_start), _ptr(self.chunk_len), c_int(self.n_chunks),
                                    _ptr(self.sq_chunk), _stream()), "kr_chunk_sqnorm")
        check(L.kr_chunk_to_tensor_sq(_ptr(self.sq_chunk), _ptr(self.first_chunk), _ptr(self.sq), c_int(self.n_tensors),
                                      _ptr(self.ctrl), _stream()), "kr_chunk_to_tensor_sq")
        check(L.kr_step_control(_ptr(self.sq), _ptr(self.t_preclip), _ptr(self.tscale), c_int(self.n_tensors),
                                _ptr(self.ctrl), c_float(clip_norm if clip_norm is not None else c.max_grad_norm),
                                _ptr(clip_override), c_float(c.adam_betas[0]), c_float(c.adam_betas[1]),
                                c_float(c.grad_explosion_abs_floor), c_float(c.grad_explosion_warmup_floor),
                                c_int(c.grad_explosion_warmup_steps), c_float(c.grad_explosion_ema_alpha),
                                c_float(c.grad_explosion_multiplier), c_int(c.grad_explosion_min_ema_steps),
                                c_float(c.emergency_clip), _stream()), "kr_step_control")
        check(L.kr_adamw_step(_ptr(s.params), _ptr(s.grads), _ptr(s.exp_avg), _ptr(s.exp_avg_sq), _ptr(s.ema),
                              _ptr(s.shadow), _ptr(self.chunk_tensor), _ptr(self.chunk_start), _ptr(self.chunk_len),
                              c_int(self.n_chunks), _ptr(self.t_group), _ptr(self.tscale), _ptr(self.g_lr),
                              _ptr(self.g_wd), _ptr(self.t_wnmax), _ptr(self.wsq), _ptr(self.ctrl),
                              c_float(c.adam_betas[0]), c_float(c.adam_betas[1]), c_float(c.adam_eps),
                              c_float(c.ema_decay), _stream()), "kr_adamw_step")
        # the step's tail: the weight-norm projection (FFN matrices) and the dgrad shadows of the predictor convs read
        # disjoint weights and both depend on the AdamW kernel only -> side by side (they were 65 serial us)
        side = None
        if self.n_wn_chunks and s.params.is_cuda:
            import torch
            if getattr(self, "_tail_stream", None) is None:
                self._tail_stream = torch.cuda.Stream(device=s.params.device)
            side, cur = self._tail_stream, torch.cuda.current_stream(s.params.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                s.refresh_conv_dgrad()
        if self.n_wn_chunks:
            check(L.kr_wn_project(_ptr(s.params), _ptr(s.shadow), _ptr(self.wn_chunks), c_int(self.n_wn_chunks),
                                  _ptr(self.chunk_tensor), _ptr(self.chunk_start), _ptr(self.chunk_len),
                                  _ptr(self.t_wnmax), _ptr(self.wsq), _ptr(self.ctrl), _stream()), "kr_wn_project")
        if side is not None:
            cur.wait_stream(side)
        else:
            s.refresh_conv_dgrad()

    def write_detector_state(self, ema_norm: float, ema_steps: int) -> None:
        """Restores the explosion detector's norm EMA (checkpoint resume)."""
        raw = self.ctrl.cpu()
        raw.view(torch.float32)[CTRL_FIELDS.index("ema_norm")] = float(ema_norm)
        raw[CTRL_FIELDS.index("ema_steps")] = int(ema_steps)
        self.ctrl.copy_(raw)

    def read_ctrl(self) -> Dict[str, float]:
        """Host copy of the device control block (synchronises; logging only)."""
        raw = self.ctrl.cpu()
        as_f = raw.view(torch.float32)
        return {k: (int(raw[i]) if k in _CTRL_INT else float(as_f[i])) for i, k in enumerate(CTRL_FIELDS)}
