"""ORACLE (test infrastructure, never the product path) — CPU restatement of the reference's autoregressive
inference, the SURVEY.md §8(f) N2 row that comes after the training step:

  KokoroModel.forward_inference          src/kokoro/model/model.py:675-779   (bounds of the generation loop)
  VarianceAdaptor.forward (inference)    model/variance_predictor.py:335-437 (durations = round(expm1(log_dur)) >= 0,
                                         predicted pitch / energy clamped to [0, 1] feed the embeddings)
  KokoroGenerator.generate               model/generator.py:24-127            (frame-by-frame decode, stop rules)
  MultiHeadAttentionImproved KV cache    model/transformers.py:237-277        (self-attention cache holds the RAW key
                                         projections and the normalised values; q/k-norm and RoPE are re-applied to the
                                         whole cache every step)

Reference quirk reproduced on purpose: RoPE is applied with q_offset = 0, so the single new query is always rotated
as position 0 while the cached keys carry positions 0..t (transformers.py:276-277, positional_encoding.py:163-209) —
training rotates query t as position t.  Parity is PINNED against the live reference
(tests/golden/make_golden_inference.py -> tests/golden/inference.npz, tests/test_inference_cpu.py).
Only tests/ may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import acoustic as oa

Tensor = torch.Tensor


def _rope_at(t: Tensor, offset: int) -> Tensor:
    """rotate-half RoPE of (B,H,S,dk) with positions offset .. offset+S-1."""
    S, dk = t.shape[-2], t.shape[-1]
    cos, sin = oa.rope_tables(offset + S, dk, t.device)
    cos, sin = cos[offset:], sin[offset:]
    half = dk // 2
    rot = torch.cat([-t[..., half:], t[..., :half]], dim=-1)
    return t * cos + rot * sin


def _heads(t: Tensor, H: int) -> Tensor:
    B, S, D = t.shape
    return t.view(B, S, H, D // H).transpose(1, 2)


def encode_and_expand(sd: Dict[str, Tensor], cfg: oa.AcousticConfig, phoneme_indices: Tensor,
                      stress_indices: Optional[Tensor], details: Optional[dict] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Inference branch of _encode_and_expand (model.py:450-508): (memory (B,T',D), frame_mask, log_dur).
    `details` (a dict) receives the intermediate predictions and selected bins (for staged parity checks)."""
    B, P = phoneme_indices.shape
    D = cfg.hidden_dim
    pe = sd["positional_encoding.pe"][0]
    va = "duration_adaptor.variance_adaptor."
    text_pad = phoneme_indices == 0
    x = sd["text_embedding.weight"][phoneme_indices] * math.sqrt(D)
    if stress_indices is not None:
        x = x + F.embedding(stress_indices, sd["stress_embedding.weight"], padding_idx=0)
    x = x + pe[:P]
    for i in range(cfg.n_encoder_layers):
        x = oa.encoder_block(sd, f"transformer_encoder_layers.{i}.", cfg, x, text_pad)
    enc = oa._ln(sd, "encoder_norm.", x)
    log_dur = oa.variance_predictor(sd, va + "duration_predictor.", cfg, enc, text_pad)
    dur = torch.clamp(torch.round(torch.expm1(log_dur)), min=0)               # variance_predictor.py:347
    mem = oa.expand_tokens(enc, dur)
    if mem.shape[1] < 3:                                                       # :357-359
        mem = F.pad(mem, (0, 0, 0, 3 - mem.shape[1]))
    Tp = mem.shape[1]
    lengths = dur.long().sum(dim=1)
    fmask = torch.arange(Tp).unsqueeze(0) >= lengths.unsqueeze(1)
    pitch = oa.variance_predictor(sd, va + "pitch_predictor.", cfg, mem, fmask)
    energy = oa.variance_predictor(sd, va + "energy_predictor.", cfg, mem, fmask)
    p_idx = torch.bucketize(pitch.clamp(0.0, 1.0), sd[va + "pitch_bins"])     # :410, :431
    e_idx = torch.bucketize(energy.clamp(0.0, 1.0), sd[va + "energy_bins"])
    if details is not None:
        details.update(pitch=pitch, energy=energy, pitch_idx=p_idx, energy_idx=e_idx, encoder=enc)
    mem = mem + sd[va + "pitch_embedding.weight"][p_idx] + sd[va + "energy_embedding.weight"][e_idx]
    return mem.masked_fill(fmask.unsqueeze(-1), 0.0), fmask, log_dur


def _self_attn_step(sd, pre: str, cfg, x: Tensor, cache: Tuple) -> Tuple[Tensor, Tuple]:
    """One new frame through the self-attention with the KV cache (transformers.py:228-277, 393-437)."""
    H, dk = cfg.n_heads, cfg.head_dim
    q = _heads(x @ sd[pre + "w_q.weight"].t(), H)
    new_k = _heads(x @ sd[pre + "w_k.weight"].t(), H)
    new_v = oa._rms(_heads(x @ sd[pre + "w_v.weight"].t(), H), sd[pre + "v_norm.weight"])
    k_raw = new_k if not cache else torch.cat([cache[0], new_k], dim=2)
    v = new_v if not cache else torch.cat([cache[1], new_v], dim=2)
    q = _rope_at(oa._rms(q, sd[pre + "q_norm.weight"]), 0)                    # q_offset = 0: the reference's quirk
    k = _rope_at(oa._rms(k_raw, sd[pre + "k_norm.weight"]), 0)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(x.shape[0], 1, -1)
    return o @ sd[pre + "w_o.weight"].t() + sd[pre + "w_o.bias"], (k_raw, v)


def _cross_attn_step(sd, pre: str, cfg, x: Tensor, k_raw: Tensor, v: Tensor, mem_pad: Tensor) -> Tensor:
    H, dk = cfg.n_heads, cfg.head_dim
    q = oa._rms(_heads(x @ sd[pre + "w_q.weight"].t(), H), sd[pre + "q_norm.weight"])
    k = oa._rms(k_raw, sd[pre + "k_norm.weight"])
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
    s = s.masked_fill(mem_pad.view(mem_pad.shape[0], 1, 1, -1), float("-inf"))
    o = (oa.softmax_rows(s) @ v).transpose(1, 2).reshape(x.shape[0], 1, -1)     # an all-padding memory gives a zero context
    return o @ sd[pre + "w_o.weight"].t() + sd[pre + "w_o.bias"]


def generation_bounds(expected: int, max_len: int = 4000, min_len_ratio: float = 0.7, min_len_floor: int = 12,
                      max_len_ratio: float = 3.0, max_len_cap: int = 1600) -> Tuple[int, int]:
    """model.py:737-745."""
    lo = max(min_len_floor, int(expected * min_len_ratio))
    hi = min(max_len, max(expected + 80, int(expected * max_len_ratio)), max_len_cap)
    if hi <= lo:
        hi = min(max_len, lo + 1)
    return lo, hi


@torch.no_grad()
def forward_inference(sd: Dict[str, Tensor], cfg: oa.AcousticConfig, phoneme_indices: Tensor,
                      stress_indices: Optional[Tensor] = None, max_len: int = 4000, stop_threshold: float = 0.5,
                      post_expected_stop_threshold: float = 0.2, return_stop_probs: bool = False,
                      return_raw: bool = False):
    """(B, n_frames, n_mels) generated mel, clamped to [-11.5, 2] (model.py:675-779 + generator.py:24-127).
    return_raw adds the un-clamped frames (what the loop feeds back) as a third result."""
    mem, mem_pad, _ = encode_and_expand(sd, cfg, phoneme_indices, stress_indices)
    B = mem.shape[0]
    expected = mem.shape[1]
    lo, hi = generation_bounds(expected, max_len)
    pe = sd["positional_encoding.pe"][0]
    H = cfg.n_heads
    cross = []
    for i in range(cfg.n_decoder_layers):                                     # precompute_cross_attention_kv
        pre = f"decoder.layers.{i}.cross_attn."
        cross.append((_heads(mem @ sd[pre + "w_k.weight"].t(), H),
                      oa._rms(_heads(mem @ sd[pre + "w_v.weight"].t(), H), sd[pre + "v_norm.weight"])))
    caches: List[Tuple] = [() for _ in range(cfg.n_decoder_layers)]
    frame = torch.zeros(B, 1, cfg.mel_dim)
    out, probs = [], []
    for t in range(hi):
        y = frame @ sd["mel_projection_in.weight"].t() + sd["mel_projection_in.bias"] + pe[t:t + 1]
        for i in range(cfg.n_decoder_layers):
            pre = f"decoder.layers.{i}."
            a, caches[i] = _self_attn_step(sd, pre + "self_attn.", cfg, oa._ln(sd, pre + "norm1.", y), caches[i])
            y = y + a
            y = y + _cross_attn_step(sd, pre + "cross_attn.", cfg, oa._ln(sd, pre + "norm2.", y), *cross[i], mem_pad)
            y = y + oa.glu_ffn(sd, pre + "ff.", oa._ln(sd, pre + "norm3.", y))
        y = oa._ln(sd, "decoder.norm.", y)
        mel_t = y @ sd["mel_projection_out.weight"].t() + sd["mel_projection_out.bias"]
        stop_t = (y @ sd["stop_token_predictor.weight"].t() + sd["stop_token_predictor.bias"]).squeeze(-1)
        out.append(mel_t)
        p = float(torch.sigmoid(stop_t).mean())
        probs.append(p)
        if t >= lo:                                                            # generator.py:66-86
            thr = stop_threshold if t < expected else min(stop_threshold, post_expected_stop_threshold)
            if p > thr:
                break
            if len(out) >= 30 and float(torch.cat(out[-30:], dim=1).mean()) < -9.5:
                break
        frame = mel_t
    raw = torch.cat(out, dim=1)
    mel = raw.clamp(min=-11.5, max=2.0)
    if return_raw:
        return mel, probs, raw
    return (mel, probs) if return_stop_probs else mel
