"""Flat parameter store for the acoustic model.

All 308 parameter tensors of the reference KokoroModel (same names, same order as
reference src/kokoro/model/model.py:35-210 registers them) live in ONE fp32 buffer, with parallel
flat buffers for the gradient, Adam moments, EMA weights and the bf16 shadow the tcgen05 GEMMs
read.  One flat gradient buffer = one NCCL all-reduce per optimizer step and one fused
multi-tensor optimizer launch.

Two layout rules differ from the reference and are hidden at the state-dict boundary:
  * w_q, w_k, w_v of an attention block are adjacent, so [w_q; w_k; w_v] is one [3D, D] GEMM
    operand (and [w_k; w_v] one [2D, D] operand for cross-attention) with no concatenation;
  * Conv1d weights are stored tap-major, [C_out, 3, C_in] instead of [C_out, C_in, 3], which is
    the K-major operand of the overlapping-row conv-as-GEMM; ``state_dict()`` / named views
    present the reference's [C_out, C_in, 3] shape through a permuted view.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

ALIGN = 64  # elements; keeps every tensor 256-byte aligned in fp32 and 128-byte aligned in bf16


@dataclass
class ModelConfig:
    """Constructor arguments of the reference KokoroModel (model/model.py:35-47) as the trainer
    passes them (training/trainer.py:356-382)."""
    vocab_size: int = 59
    mel_dim: int = 80
    hidden_dim: int = 512
    n_encoder_layers: int = 6
    n_heads: int = 8
    encoder_ff_dim: int = 1536
    n_decoder_layers: int = 6
    decoder_ff_dim: int = 1536
    max_decoder_seq_len: int = 4000
    variance_filter_size: int = 256
    variance_kernel_size: int = 3
    n_variance_bins: int = 256
    use_stress_embedding: bool = True
    qk_norm: bool = True
    ffn_output_norm: bool = True
    vp_chunk: int = 512


def param_specs(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, reference shape) of every parameter, in the reference's registration order."""
    D, F, V, ff_e, ff_d = cfg.hidden_dim, cfg.variance_filter_size, cfg.vocab_size, cfg.encoder_ff_dim, cfg.decoder_ff_dim
    dk = D // cfg.n_heads
    out: List[Tuple[str, Tuple[int, ...]]] = [("text_embedding.weight", (V, D)), ("stress_embedding.weight", (3, D))]

    def attn(p):
        return [(p + "w_q.weight", (D, D)), (p + "w_k.weight", (D, D)), (p + "w_v.weight", (D, D)),
                (p + "w_o.weight", (D, D)), (p + "w_o.bias", (D,)), (p + "q_norm.weight", (dk,)),
                (p + "k_norm.weight", (dk,)), (p + "v_norm.weight", (dk,))]

    def ffn(p, ff):
        return [(p + "linear1.weight", (2 * ff, D)), (p + "linear1.bias", (2 * ff,)),
                (p + "linear2.weight", (D, ff)), (p + "linear2.bias", (D,)), (p + "output_norm.weight", (D,))]

    def ln(p):
        return [(p + "weight", (D,)), (p + "bias", (D,))]

    for i in range(cfg.n_encoder_layers):
        p = f"transformer_encoder_layers.{i}."
        out += attn(p + "self_attn.") + ffn(p + "ff.", ff_e) + ln(p + "norm1.") + ln(p + "norm2.")
    out += ln("encoder_norm.")
    va = "duration_adaptor.variance_adaptor."
    for name in ("duration_predictor.", "pitch_predictor.", "energy_predictor."):
        p = va + name
        out += [(p + "conv_layers.0.weight", (F, D, 3)), (p + "conv_layers.0.bias", (F,)),
                (p + "conv_layers.1.weight", (F, F, 3)), (p + "conv_layers.1.bias", (F,)),
                (p + "norms.0.weight", (F,)), (p + "norms.0.bias", (F,)),
                (p + "norms.1.weight", (F,)), (p + "norms.1.bias", (F,)),
                (p + "linear.weight", (1, F)), (p + "linear.bias", (1,))]
    out += [(va + "pitch_embedding.weight", (cfg.n_variance_bins, D)),
            (va + "energy_embedding.weight", (cfg.n_variance_bins, D)),
            ("mel_projection_in.weight", (D, cfg.mel_dim)), ("mel_projection_in.bias", (D,))]
    for i in range(cfg.n_decoder_layers):
        p = f"decoder.layers.{i}."
        out += (attn(p + "self_attn.") + attn(p + "cross_attn.") + ffn(p + "ff.", ff_d) + ln(p + "norm1.")
                + ln(p + "norm2.") + ln(p + "norm3."))
    out += ln("decoder.norm.")
    out += [("mel_projection_out.weight", (cfg.mel_dim, D)), ("mel_projection_out.bias", (cfg.mel_dim,)),
            ("stop_token_predictor.weight", (1, D)), ("stop_token_predictor.bias", (1,))]
    return out


BUFFER_NAMES = ("positional_encoding.pe", "duration_adaptor.variance_adaptor.pitch_bins",
                "duration_adaptor.variance_adaptor.energy_bins")


def sinusoid_table(max_len: int, dim: int) -> torch.Tensor:
    """positional_encoding.pe buffer (reference model/positional_encoding.py:20-34)."""
    pos = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros(max_len, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def rope_tables(max_len: int, dk: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin [max_len, dk/2] (reference model/positional_encoding.py:96-160, base 1e4)."""
    theta = 1.0 / (10000.0 ** (torch.arange(0, dk, 2, dtype=torch.float32) / dk))
    ang = torch.outer(torch.arange(max_len, dtype=torch.float32), theta)
    return ang.cos().contiguous(), ang.sin().contiguous()


def _is_conv(name: str) -> bool:
    return ".conv_layers." in name and name.endswith(".weight")


@dataclass
class _Entry:
    name: str
    shape: Tuple[int, ...]       # reference shape
    offset: int
    numel: int


class ParamStore:
    """Flat fp32 master / grad / moment / EMA buffers + bf16 shadow, with named views."""

    def __init__(self, cfg: ModelConfig, device: torch.device, with_ema: bool = True):
        self.cfg = cfg
        self.device = device
        self.entries: Dict[str, _Entry] = {}
        self.order: List[str] = []
        off = 0
        for name, shape in param_specs(cfg):
            n = 1
            for s in shape:
                n *= s
            self.entries[name] = _Entry(name, shape, off, n)
            self.order.append(name)
            off += (n + ALIGN - 1) // ALIGN * ALIGN
        self.total = off
        self._views: Dict[Tuple[int, str], torch.Tensor] = {}
        self.params = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=device)
        self.ema = torch.zeros(off, dtype=torch.float32, device=device) if with_ema else None
        self.shadow = torch.zeros(off, dtype=torch.bfloat16, device=device)
        # buffers of the reference state dict
        self.pe = sinusoid_table(cfg.max_decoder_seq_len, cfg.hidden_dim).to(device)
        self.pitch_bins = torch.linspace(0.0, 1.0, cfg.n_variance_bins - 1).to(device)
        self.energy_bins = torch.linspace(0.0, 1.0, cfg.n_variance_bins - 1).to(device)
        cos, sin = rope_tables(cfg.max_decoder_seq_len, cfg.hidden_dim // cfg.n_heads)
        self.rope_cos, self.rope_sin = cos.to(device), sin.to(device)
        # dgrad shadows of the conv weights that need a data gradient
        self.conv_dgrad: Dict[str, torch.Tensor] = {}
        for name in self.order:
            if _is_conv(name) and self._conv_needs_dgrad(name):
                co, ci, _ = self.entries[name].shape
                self.conv_dgrad[name] = torch.zeros(ci, 3 * co, dtype=torch.bfloat16, device=device)

    @staticmethod
    def _conv_needs_dgrad(name: str) -> bool:
        # conv_layers.1 always; conv_layers.0 only for the duration predictor (the pitch / energy
        # predictors read the detached expansion, reference utils/lengths.py:30)
        return ".conv_layers.1." in name or "duration_predictor.conv_layers.0." in name

    # ----- views -------------------------------------------------------------------------------
    def _flat(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        e = self.entries[name]
        return buf[e.offset:e.offset + e.numel]

    def _internal_shape(self, name: str) -> Tuple[int, ...]:
        e = self.entries[name]
        if _is_conv(name):
            co, ci, k = e.shape
            return (co, k * ci)
        return e.shape

    def _view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        """Cached internal-layout view of `name` inside flat buffer `buf`.  An eager step asks for ~400 of these (two torch
        calls each: ~1.5 ms of the ~17 ms a host-launch-bound eager step takes); the key carries the buffer's address because
        the buffers are re-pointed (EMA weights for validation, symmetric-memory gradients under data parallelism)."""
        key = (buf.data_ptr(), name)
        v = self._views.get(key)
        if v is None:
            v = self._views[key] = self._flat(buf, name).view(self._internal_shape(name))
        return v

    def p(self, name: str) -> torch.Tensor:
        """fp32 master weight in the INTERNAL layout (conv: [C_out, 3*C_in])."""
        return self._view(self.params, name)

    def g(self, name: str) -> torch.Tensor:
        return self._view(self.grads, name)

    def w(self, name: str) -> torch.Tensor:
        """bf16 shadow in the internal layout."""
        return self._view(self.shadow, name)

    def span(self, buf: torch.Tensor, first: str, rows: int, cols: int) -> torch.Tensor:
        """[rows, cols] view over consecutive tensors starting at `first` (fused QKV / KV)."""
        e = self.entries[first]
        return buf[e.offset:e.offset + rows * cols].view(rows, cols)

    def ref_view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        """View with the REFERENCE shape (conv weights permuted back to [C_out, C_in, 3])."""
        e = self.entries[name]
        flat = self._flat(buf, name)
        if _is_conv(name):
            co, ci, k = e.shape
            return flat.view(co, k, ci).permute(0, 2, 1)
        return flat.view(e.shape)

    # ----- state dict boundary -----------------------------------------------------------------
    def state_dict(self, buf: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        buf = self.params if buf is None else buf
        sd: Dict[str, torch.Tensor] = {}
        for name in self.order:
            sd[name] = self.ref_view(buf, name)
        sd["positional_encoding.pe"] = self.pe.unsqueeze(0)
        sd["duration_adaptor.variance_adaptor.pitch_bins"] = self.pitch_bins
        sd["duration_adaptor.variance_adaptor.energy_bins"] = self.energy_bins
        return sd

    def ordered_state_dict(self) -> Dict[str, torch.Tensor]:
        """state dict in the reference's key order (buffers interleaved where the reference has them)."""
        sd = self.state_dict()
        out: Dict[str, torch.Tensor] = {}
        for name in self.order:
            if name == "transformer_encoder_layers.0.self_attn.w_q.weight":
                out["positional_encoding.pe"] = sd["positional_encoding.pe"]
            if name == "duration_adaptor.variance_adaptor.duration_predictor.conv_layers.0.weight":
                out[BUFFER_NAMES[1]] = sd[BUFFER_NAMES[1]]
                out[BUFFER_NAMES[2]] = sd[BUFFER_NAMES[2]]
            out[name] = sd[name]
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        missing = [n for n in self.order if n not in sd]
        unexpected = [k for k in sd if k not in self.entries and k not in BUFFER_NAMES]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing={missing[:5]} unexpected={unexpected[:5]}")
        with torch.no_grad():
            for name in self.order:
                if name in sd:
                    src = sd[name].to(device=self.device, dtype=torch.float32)
                    if tuple(src.shape) != tuple(self.entries[name].shape):
                        raise RuntimeError(f"shape mismatch for {name}: {tuple(src.shape)} vs {self.entries[name].shape}")
                    self.ref_view(self.params, name).copy_(src)
            for key, attr in zip(BUFFER_NAMES, ("pe", "pitch_bins", "energy_bins")):
                if key in sd:
                    t = sd[key].to(device=self.device, dtype=torch.float32)
                    getattr(self, attr).copy_(t.reshape(getattr(self, attr).shape))
        self.refresh_shadow()
        if self.ema is not None:
            self.ema.copy_(self.params)

    def refresh_shadow(self) -> None:
        """bf16 shadow <- fp32 master (whole buffer) and the conv dgrad shadows."""
        from . import ops
        ops.cast_bf16(self.params, self.shadow)
        self.refresh_conv_dgrad()

    def refresh_conv_dgrad(self) -> None:
        """bf16 dgrad layouts of the variance predictors' conv weights: one launch per (Co, Ci) shape."""
        from . import ops
        by_shape: Dict[Tuple[int, int], list] = {}
        for name, wd in self.conv_dgrad.items():
            co, ci, _ = self.entries[name].shape
            by_shape.setdefault((co, ci), []).append((self.p(name), wd))
        for (co, ci), pairs in by_shape.items():
            for k in range(0, len(pairs), 8):
                ops.conv_dgrad_shadow_multi(pairs[k:k + 8], co, ci)

    # ----- default initialisation (reference nn.Module defaults + explicit inits) ---------------
    def init_default(self, seed: int = 0) -> None:
        """Same distributions as the reference modules' defaults (nn.Linear / nn.Conv1d kaiming-uniform
        a=sqrt(5), nn.Embedding N(0,1), LayerNorm/RMSNorm/GroupNorm ones/zeros) plus the reference's
        explicit inits (model/model.py:82-86,169-186; variance_predictor duration bias log1p(5)).
        The random stream differs from torch.manual_seed(seed)+reference construction order."""
        g = torch.Generator().manual_seed(seed)
        D = self.cfg.hidden_dim
        sd: Dict[str, torch.Tensor] = {}
        specs = dict(param_specs(self.cfg))
        for name, shape in specs.items():
            leaf = name.rsplit(".", 1)[-1]
            if name == "text_embedding.weight":
                t = torch.randn(shape, generator=g) / math.sqrt(D)
            elif name == "stress_embedding.weight":
                t = torch.randn(shape, generator=g)
                t[0].zero_()                      # padding_idx=0
            elif name.endswith("_embedding.weight"):
                t = torch.randn(shape, generator=g)
            elif leaf == "weight" and len(shape) == 1:
                t = torch.ones(shape)
            elif leaf == "bias":
                wname = name[:-4] + "weight"
                wshape = specs.get(wname, None)
                if wshape is None or len(wshape) == 1 or name.startswith("mel_projection") or name.startswith("stop_token"):
                    t = torch.zeros(shape)
                else:
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    bound = 1.0 / math.sqrt(fan_in)
                    t = (torch.rand(shape, generator=g) * 2 - 1) * bound
                if name.endswith("duration_predictor.linear.bias"):
                    t = torch.full(shape, math.log1p(5.0))
            elif name in ("mel_projection_in.weight", "mel_projection_out.weight", "stop_token_predictor.weight"):
                bound = math.sqrt(6.0 / (shape[0] + shape[1]))   # xavier_uniform
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                bound = 1.0 / math.sqrt(fan_in)   # kaiming_uniform(a=sqrt(5))
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            sd[name] = t
        self.load_state_dict(sd, strict=False)
