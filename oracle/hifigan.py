"""ORACLE (test infrastructure, never the product path) — CPU restatement of HiFi-GAN v1 generator
inference of igorshmukler/kokoro-ruslan in plain fp32 torch.

Follows reference src/kokoro/inference/hifigan_vocoder.py: ResBlock.forward :63-70,
HiFiGANGenerator.forward :110-133 (input layout handling :112-117, final leaky_relu with the DEFAULT
slope 0.01 :130), weight_norm parametrisation g * v / ||v|| with the norm over all dims but dim 0
(= per output channel for Conv1d, per INPUT channel for ConvTranspose1d).  Functional over a flat
state dict with the reference's keys (``*.parametrizations.weight.original0`` = g, ``original1`` = v).
Parity is PINNED: tests/golden/hifigan_small.npz is generated from the live reference module by
tests/golden/make_golden_hifigan.py.  Only tests/, smoke() and bench.py's CPU legs import this.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class HifiConfig:
    """vocoder_models/hifigan/config_universal_v1.json"""
    upsample_rates: Tuple[int, ...] = (8, 8, 2, 2)
    upsample_kernel_sizes: Tuple[int, ...] = (16, 16, 4, 4)
    upsample_initial_channel: int = 512
    resblock_kernel_sizes: Tuple[int, ...] = (3, 7, 11)
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    num_mels: int = 80


def state_dict_keys(cfg: HifiConfig) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) in the reference module's registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def wn(prefix, wshape):
        g = (wshape[0],) + (1,) * (len(wshape) - 1)
        return [(prefix + ".bias", None), (prefix + ".parametrizations.weight.original0", g),
                (prefix + ".parametrizations.weight.original1", wshape)]

    c0 = cfg.upsample_initial_channel
    for k, s in wn("conv_pre", (c0, cfg.num_mels, 7)):
        out.append((k, s if s is not None else (c0,)))
    for i, (u, ks) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = c0 // 2 ** i, c0 // 2 ** (i + 1)
        for k, s in wn(f"ups.{i}", (cin, cout, ks)):
            out.append((k, s if s is not None else (cout,)))
    nk = len(cfg.resblock_kernel_sizes)
    ch = c0
    for i in range(len(cfg.upsample_rates)):
        ch = c0 // 2 ** (i + 1)
        for j, (ks, dil) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            for group in ("convs1", "convs2"):
                for d in range(len(dil)):
                    for k, s in wn(f"resblocks.{i * nk + j}.{group}.{d}", (ch, ch, ks)):
                        out.append((k, s if s is not None else (ch,)))
    for k, s in wn("conv_post", (1, ch, 7)):
        out.append((k, s if s is not None else (1,)))
    return out


def seeded_state_dict(cfg: HifiConfig, seed: int = 0, gain_jitter: float = 0.3) -> Dict[str, Tensor]:
    """Deterministic weights: v ~ N(0, 1/fan_in), g = ||v|| * (1 + jitter * N(0,1)) (so the weight-norm
    gain is exercised away from its initial value), biases ~ 0.1 N(0,1)."""
    sd: Dict[str, Tensor] = {}
    keys = state_dict_keys(cfg)
    for i, (name, shape) in enumerate(keys):
        g = torch.Generator().manual_seed(seed * 7919 + i)
        if name.endswith(".bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("original1"):
            fan = 1
            for s in shape[1:]:
                fan *= s
            if name.startswith("ups."):
                fan = shape[0] * shape[2] // max(1, cfg.upsample_rates[int(name.split(".")[1])])
            scale = 1.0 / fan ** 0.5
            if ".convs2." in name:
                scale *= 0.35          # keeps the residual branches from blowing the activations up
            if name.startswith("conv_post"):
                scale *= 0.25          # pre-tanh output ~O(0.5): audio is not saturated
            sd[name] = torch.randn(shape, generator=g) * scale
    for name, shape in keys:
        if name.endswith("original0"):
            v = sd[name[:-1] + "1"]
            g = torch.Generator().manual_seed(seed * 7919 + 100000 + len(name))
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
            sd[name] = nrm * (1.0 + gain_jitter * torch.randn(shape, generator=g)).abs()
    return {k: sd[k] for k, _ in keys}


def effective_weight(sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """W = g * v / ||v||, norm over all dims except dim 0 (torch.nn.utils.parametrizations.weight_norm
    default dim=0; hifigan_vocoder.py:37-102)."""
    g = sd[prefix + ".parametrizations.weight.original0"]
    v = sd[prefix + ".parametrizations.weight.original1"]
    nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / nrm)


def normalize_input(x: Tensor) -> Tensor:
    """(B,80,T) passthrough; (B,T,80) and (T,80) transposed (hifigan_vocoder.py:112-117)."""
    if x.dim() == 3 and x.size(1) != 80 and x.size(2) == 80:
        x = x.transpose(1, 2)
    elif x.dim() == 2:
        x = x.unsqueeze(0).transpose(1, 2)
    return x


def generator_forward(sd: Dict[str, Tensor], cfg: HifiConfig, x: Tensor) -> Tensor:
    x = normalize_input(x).float()
    h = F.conv1d(x, effective_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    nk = len(cfg.resblock_kernel_sizes)
    for i, (u, ks) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        h = F.leaky_relu(h, 0.1)
        h = F.conv_transpose1d(h, effective_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u,
                               padding=(ks - u) // 2)
        xs = None
        for j, (k, dils) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            p = f"resblocks.{i * nk + j}"
            y = h
            for d, dil in enumerate(dils):
                t = F.leaky_relu(y, 0.1)
                t = F.conv1d(t, effective_weight(sd, f"{p}.convs1.{d}"), sd[f"{p}.convs1.{d}.bias"],
                             padding=(k * dil - dil) // 2, dilation=dil)
                t = F.leaky_relu(t, 0.1)
                t = F.conv1d(t, effective_weight(sd, f"{p}.convs2.{d}"), sd[f"{p}.convs2.{d}.bias"],
                             padding=(k - 1) // 2)
                y = t + y
            xs = y if xs is None else xs + y
        h = xs / nk
    h = F.leaky_relu(h)            # default slope 0.01 (hifigan_vocoder.py:130)
    h = F.conv1d(h, effective_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    return torch.tanh(h)


def synthetic_mel(B: int = 16, T: int = 800, seed: int = 0) -> Tensor:
    """SURVEY.md §8(d) config-5 input: N(0,1)*2-5, (B,80,T)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 80, T, generator=g) * 2.0 - 5.0
