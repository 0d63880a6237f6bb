"""HiFi-GAN v1 generator inference on B200 — host-side mirror of the reference interface
(src/kokoro/inference/hifigan_vocoder.py: ``AttrDict`` :24-28, ``HiFiGANGenerator`` :79-141,
``HiFiGANConfig`` :144-187, ``load_hifigan_model`` :190-271): same constructor argument, same
``forward`` input layouts, same state-dict keys (``*.parametrizations.weight.original0/1``).

Device side: every Conv1d / ConvTranspose1d is ONE launch of the tcgen05 implicit-GEMM kernel
(kr_gemm_ex, conv mode) on channels-last bf16 activations [B, L + 2*HALO, C] whose zero halos double
as the convolution padding, with the elementwise work fused into the epilogues:

    conv_pre      -> lrelu(0.1)                                   (C2 output only)
    ups[i]        -> polyphase ConvTranspose as a 3-tap conv with N = stride*C_out; writes h and lrelu(h)
    ResBlock c1   -> lrelu(conv + bias)                           (C2 only)
    ResBlock c2   -> y' = y + conv + bias;  writes y' (fp32 residual stream) and lrelu(y') (bf16 operand)
    last c2 of resblock j -> xs (+)= y'/3 (fp32 MRF accumulator); the third one writes lrelu(xs) for the
                     next stage (slope 0.1, or the reference's default 0.01 before conv_post :130)
    conv_post+tanh-> kr_hifi_post_tanh (N = 1: HBM-bound GEMV)

weight_norm (g * v / ||v||) is folded once per weight update, not per forward as in the reference.
The whole forward (79 launches) is captured in a CUDA graph per (B, T).
"""
from __future__ import annotations

import ctypes
import json
import math
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from ._lib import check, lib, launch_count

HALO = 32          # >= max dilation * (k-1)/2 = 5*5 = 25 rows of zero padding on each side
BF16, F32 = torch.bfloat16, torch.float32


class AttrDict(dict):
    """Dictionary that allows attribute access (reference hifigan_vocoder.py:24-28)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class HiFiGANConfig:
    @staticmethod
    def get_default_config() -> AttrDict:
        """vocoder_models/hifigan/config_universal_v1.json (reference :148-176)."""
        return AttrDict({
            "resblock": "1", "num_gpus": 0, "batch_size": 16, "learning_rate": 0.0002, "adam_b1": 0.8,
            "adam_b2": 0.99, "lr_decay": 0.999, "seed": 1234, "upsample_rates": [8, 8, 2, 2],
            "upsample_kernel_sizes": [16, 16, 4, 4], "upsample_initial_channel": 512,
            "resblock_kernel_sizes": [3, 7, 11], "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
            "segment_size": 8192, "num_mels": 80, "num_freq": 1025, "n_fft": 1024, "hop_size": 256,
            "win_size": 1024, "sampling_rate": 22050, "fmin": 0, "fmax": 8000, "fmax_for_loss": None})

    @staticmethod
    def load_config(config_path) -> AttrDict:
        try:
            with open(config_path, "r") as f:
                return AttrDict(json.load(f))
        except Exception:
            return HiFiGANConfig.get_default_config()


def _cpad(c: int) -> int:
    """Physical channel count of an activation: the implicit-GEMM K block is 64 channels (128-byte swizzle),
    or exactly 32 channels (64-byte swizzled K blocks, last HiFi-GAN stage)."""
    return 32 if c == 32 else (c + 63) // 64 * 64


class _Plan:
    """Buffers + captured graph for one (B, T)."""

    def __init__(self):
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.bufs: Dict[str, torch.Tensor] = {}
        self.launches = 0
        self.warm = False


class HiFiGANGenerator:
    def __init__(self, h: AttrDict, device="cuda", use_graphs: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("HiFiGANGenerator needs a CUDA device: the vocoder path has no CPU fallback")
        self.h = h
        self.device = torch.device(device)
        self.use_graphs = use_graphs
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        self.n_mels = int(getattr(h, "num_mels", 80))
        c0 = h.upsample_initial_channel
        for k, dils in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
            if max(dils) * (k - 1) // 2 > HALO:
                raise RuntimeError("resblock receptive field exceeds the activation halo")
        for u, k in zip(h.upsample_rates, h.upsample_kernel_sizes):
            if k != 2 * u:
                raise RuntimeError("polyphase ConvTranspose1d needs kernel_size == 2 * stride (HiFi-GAN v1)")
        # ---- parameters in the reference's layout / key order -------------------------------------
        self._shapes: List[Tuple[str, Tuple[int, ...]]] = []

        def wn(prefix, wshape, bias_n):
            g = (wshape[0],) + (1,) * (len(wshape) - 1)
            self._shapes += [(prefix + ".bias", (bias_n,)), (prefix + ".parametrizations.weight.original0", g),
                             (prefix + ".parametrizations.weight.original1", wshape)]
        wn("conv_pre", (c0, self.n_mels, 7), c0)
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            wn(f"ups.{i}", (c0 // 2 ** i, c0 // 2 ** (i + 1), k), c0 // 2 ** (i + 1))
        ch = c0
        for i in range(self.num_upsamples):
            ch = c0 // 2 ** (i + 1)
            for j, (k, d) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                for grp in ("convs1", "convs2"):
                    for di in range(len(d)):
                        wn(f"resblocks.{i * self.num_kernels + j}.{grp}.{di}", (ch, ch, k), ch)
        wn("conv_post", (1, ch, 7), 1)
        self._sd: Dict[str, torch.Tensor] = {}
        self._init_default(seed=int(getattr(h, "seed", 1234)))
        self._folded = False
        self._w: Dict[str, torch.Tensor] = {}
        self._plans: Dict[Tuple[int, int, bool], _Plan] = {}
        self.launches_last_forward = 0

    # ---------------------------------------------------------------------------------------------
    # parameters
    # ---------------------------------------------------------------------------------------------
    def _init_default(self, seed: int) -> None:
        """Reference initialisation: N(0, 0.01) on ups / resblocks / conv_post weights (:56-58,:106-108),
        torch's default kaiming-uniform for conv_pre, g = ||v|| (weight_norm)."""
        g = torch.Generator().manual_seed(seed)
        for name, shape in self._shapes:
            if name.endswith(".bias"):
                wshape = dict(self._shapes)[name[:-5] + ".parametrizations.weight.original1"]
                fan_in = wshape[1] * wshape[2]
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
            elif name.endswith("original1"):
                if name.startswith("conv_pre"):
                    bound = 1.0 / math.sqrt(shape[1] * shape[2])
                    t = (torch.rand(shape, generator=g) * 2 - 1) * bound
                else:
                    t = torch.randn(shape, generator=g) * 0.01
            else:
                continue
            self._sd[name] = t.to(self.device)
        for name, shape in self._shapes:
            if name.endswith("original0"):
                v = self._sd[name[:-1] + "1"]
                self._sd[name] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
        self._sd = {k: self._sd[k] for k, _ in self._shapes}

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self._sd)

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        missing = [k for k, _ in self._shapes if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for HiFiGANGenerator: missing {missing[:4]}, "
                               f"unexpected {unexpected[:4]} (weight_g / weight_v style keys are not accepted)")
        for k, shape in self._shapes:
            if k in sd:
                t = sd[k].detach().to(self.device, F32)
                if tuple(t.shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(t.shape)} vs {tuple(shape)}")
                self._sd[k] = t.contiguous()
        self._folded = False
        return self

    def named_parameters(self):
        return list(self._sd.items())

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("HiFiGANGenerator is CUDA-only (no CPU fallback)")
        return self

    def remove_weight_norm(self):
        """The reference strips the parametrisation in place; here the folded weights are what the
        kernels read anyway, so this only (re)folds."""
        self._fold()

    # ---------------------------------------------------------------------------------------------
    # weight folding (setup, runs when the weights change — not per forward)
    # ---------------------------------------------------------------------------------------------
    def _eff(self, prefix: str) -> torch.Tensor:
        g = self._sd[prefix + ".parametrizations.weight.original0"]
        v = self._sd[prefix + ".parametrizations.weight.original1"]
        nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
        return v * (g / nrm)

    @staticmethod
    def _tap_major(w: torch.Tensor, cin_phys: int, rows_phys: Optional[int] = None) -> torch.Tensor:
        """[C_out, C_in, k] -> bf16 [rows, k * cin_phys] (tap-major K, zero padded channels / rows)."""
        co, ci, k = w.shape
        rows = rows_phys or co
        out = torch.zeros(rows, k, cin_phys, dtype=F32, device=w.device)
        out[:co, :, :ci] = w.permute(0, 2, 1)
        return out.reshape(rows, k * cin_phys).to(BF16).contiguous()

    @staticmethod
    def _folded_taps(k: int, dil: int = 1):
        """Folded tap offsets f (ascending) with a non-zero block: some o = 2f + c - a (a, c in {0, 1}) is a tap offset
        j * dil, |j| <= (k-1)/2, of the original conv."""
        hh = (k - 1) // 2
        reach = (hh * dil + 1) // 2
        orig = {j * dil for j in range(-hh, hh + 1)}
        return [f for f in range(-reach, reach + 1) if any((2 * f + c - a) in orig for a in (0, 1) for c in (0, 1))]

    @classmethod
    def _tap_major_time_folded(cls, w: torch.Tensor, dil: int = 1) -> torch.Tensor:
        """Conv [C, C, k] (dilation dil, odd) on a [L, C] activation as a conv on its TIME-FOLDED view [L/2, 2C] (two
        consecutive time steps side by side in the channel dimension — the same memory): output row r holds y(2r) | y(2r+1),
        folded tap f reads x(2(r+f)) | x(2(r+f)+1), so block (row half a, column half c) of tap f is the original tap at time
        offset o = 2f + c - a.  With dil = 1, k taps become 2*ceil(h/2) + 1 (h = (k-1)/2) contiguous taps of twice the
        width; a dilated conv becomes the NON-equidistant taps of _folded_taps (only the fused ResBlock kernel takes those).
        The MMA work grows (half or more of each block-sparse weight is zero) but the narrow stage is bound by per-TILE
        latency, and the tile count halves.  Returns bf16 [2C, taps * 2C] (tap-major K)."""
        co, ci, k = w.shape
        hh = (k - 1) // 2
        taps = cls._folded_taps(k, dil)
        out = torch.zeros(2 * co, len(taps), 2 * ci, dtype=F32, device=w.device)
        for fi, f in enumerate(taps):
            for a in (0, 1):
                for c in (0, 1):
                    o = 2 * f + c - a
                    if o % dil == 0 and abs(o // dil) <= hh:
                        out[a * co:(a + 1) * co, fi, c * ci:(c + 1) * ci] = w[:, :, o // dil + hh]
        return out.reshape(2 * co, len(taps) * 2 * ci).to(BF16).contiguous()

    @classmethod
    def _half_blocks(cls, w: torch.Tensor, dil: int, folded: bool):
        """Operand of the fused ResBlock kernel (csrc/kr_hifi_resblock.cu): the conv as a list of K-HALF BLOCKS, block i =
        [64 output channels x 32 input channels] applied to the activation row at offset off[i] and input-channel half
        kh[i].  A plain 64-channel conv [64, 64, k]: two blocks per tap at offset (tau - h) * dil.  folded = True: a
        32-channel conv [32, 32, k] on the TIME-FOLDED view [L/2, 64] (see _tap_major_time_folded): block (f, c) holds, for
        output half a (rows a*32 .. a*32+31), the original tap at time offset o = 2f + c - a where that is a tap of the conv,
        zeros otherwise — blocks that are entirely zero are not listed (half of them for dilation 1, ~3/4 when dilated).
        Returns (bf16 [64, n * 32], offsets, halves)."""
        co, ci, k = w.shape
        hh = (k - 1) // 2
        blocks, offs, khs = [], [], []
        if not folded:
            assert co == 64 and ci == 64
            for tau in range(k):
                for c in (0, 1):
                    blocks.append(w[:, c * 32:(c + 1) * 32, tau])
                    offs.append((tau - hh) * dil)
                    khs.append(c)
        else:
            assert co == 32 and ci == 32
            for f in cls._folded_taps(k, dil):
                for c in (0, 1):
                    blk = torch.zeros(64, 32, dtype=w.dtype, device=w.device)
                    used = False
                    for a in (0, 1):
                        o = 2 * f + c - a
                        if o % dil == 0 and abs(o // dil) <= hh:
                            blk[a * 32:(a + 1) * 32] = w[:, :, o // dil + hh]
                            used = True
                    if used:
                        blocks.append(blk)
                        offs.append(f)
                        khs.append(c)
        return torch.cat(blocks, dim=1).to(BF16).contiguous(), offs, khs

    def _time_folded(self, stage: int) -> bool:
        """Stages whose dilation-1 convs run on the time-folded view (see _tap_major_time_folded): the 32-channel stage
        (measured on B200, 16 x 800 frames: 18.59 -> 17.67 ms; folding the 64-channel stage as well gives 17.67 ms for its
        k = 3 convs and 17.9 - 18.0 ms with k = 7 / 11, whose folded weights no longer fit next to the activation ring)."""
        return self.h.upsample_initial_channel // 2 ** (stage + 1) == 32

    def _fold(self) -> None:
        h, W = self.h, {}
        c0 = h.upsample_initial_channel
        W["pre.w"] = self._tap_major(self._eff("conv_pre"), _cpad(self.n_mels))
        W["pre.b"] = self._sd["conv_pre.bias"].contiguous()
        for i, (s, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin, cout = c0 // 2 ** i, c0 // 2 ** (i + 1)
            cin_p, cout_p = _cpad(cin), _cpad(cout)
            w = self._eff(f"ups.{i}")                         # [cin, cout, k]
            pad = (k - s) // 2
            wp = torch.zeros(s, cout_p, 3, cin_p, dtype=F32, device=w.device)
            for r in range(s):
                for j in range(3):                            # tap j reads input q - 1 + j
                    kk = s * (1 - j) + r + pad
                    if 0 <= kk < k:
                        wp[r, :cout, j, :cin] = w[:, :, kk].t()
            W[f"up{i}.w"] = wp.reshape(s * cout_p, 3 * cin_p).to(BF16).contiguous()
            bp = torch.zeros(s, cout_p, dtype=F32, device=w.device)
            bp[:, :cout] = self._sd[f"ups.{i}.bias"]
            W[f"up{i}.b"] = bp.reshape(-1).contiguous()
            for j in range(self.num_kernels):
                p = f"resblocks.{i * self.num_kernels + j}"
                for grp in ("convs1", "convs2"):
                    for d in range(len(h.resblock_dilation_sizes[j])):
                        W[f"{p}.{grp}.{d}.w"] = self._tap_major(self._eff(f"{p}.{grp}.{d}"), cout_p)
                        W[f"{p}.{grp}.{d}.b"] = self._sd[f"{p}.{grp}.{d}.bias"].contiguous()
                        dil = 1 if grp == "convs2" else h.resblock_dilation_sizes[j][d]
                        if self._time_folded(i) and dil == 1 and cout_p == cout:
                            W[f"{p}.{grp}.{d}.w2"] = self._tap_major_time_folded(self._eff(f"{p}.{grp}.{d}"))
                            W[f"{p}.{grp}.{d}.b2"] = torch.cat([W[f"{p}.{grp}.{d}.b"]] * 2).contiguous()
                # fused ResBlock step (csrc/kr_hifi_resblock.cu): 64 physical channels — the 64-channel stage as it is, the
                # 32-channel stage time-folded — whenever the half-block weights of both convs stay resident in shared memory
                folded = self._time_folded(i) and cout_p == cout
                if cout_p == 64 and cout == 64 or folded:
                    for d, dil in enumerate(h.resblock_dilation_sizes[j]):
                        w1, o1, k1 = self._half_blocks(self._eff(f"{p}.convs1.{d}"), dil, folded)
                        w2, o2, k2 = self._half_blocks(self._eff(f"{p}.convs2.{d}"), 1, folded)
                        if lib().kr_hifi_resblock_resident(ctypes.c_int(len(o1)), ctypes.c_int(len(o2))):
                            rep_b = (lambda t: torch.cat([t] * 2).contiguous()) if folded else (lambda t: t)
                            W[f"{p}.rb.{d}"] = dict(w1=w1, off1=o1, kh1=k1, b1=rep_b(W[f"{p}.convs1.{d}.b"]), w2=w2, off2=o2, kh2=k2,
                                                    b2=rep_b(W[f"{p}.convs2.{d}.b"]), folded=folded)
        wpost = self._eff("conv_post")                        # [1, ch, 7]
        W["post.w"] = wpost[0].t().contiguous()               # [7, ch] fp32
        W["post.b"] = self._sd["conv_post.bias"].contiguous()
        self._w = W
        self._folded = True
        for pl in self._plans.values():                       # captured graphs hold the old weight pointers
            pl.graph = None

    # ---------------------------------------------------------------------------------------------
    # forward
    # ---------------------------------------------------------------------------------------------
    def _alloc(self, plan: _Plan, B: int, T: int) -> None:
        h = self.h
        c0 = h.upsample_initial_channel
        dev = self.device
        z = lambda *s, dt=BF16: torch.zeros(*s, dtype=dt, device=dev)   # noqa: E731
        b = plan.bufs
        b["mel_cl"] = z(B, T + 2 * HALO, _cpad(self.n_mels))
        b["x_act"] = z(B, T + 2 * HALO, _cpad(c0))            # lrelu(conv_pre) = input of ups[0]
        L = T
        for i, s in enumerate(h.upsample_rates):
            L *= s
            cp = _cpad(c0 // 2 ** (i + 1))
            for name in ("h_act", "ya_act", "yb_act", "t_act", "x_act"):
                b[f"{name}{i}"] = z(B, L + 2 * HALO, cp)
            # residual stream in fp32: bf16 here is the dominant error term (measured 1.0-1.3e-2 on the
            # audio vs 0.75e-2 with fp32), the conv operands (the *_act tensors) stay bf16.  Re-measured at the end of
            # round 2 per stage (2 x 800 frames, gate 1e-2): fp32 everywhere 0.91e-2; bf16 in the 256-channel stage only
            # 0.94e-2; in the 256- and 128-channel stages 1.07e-2 — and neither was faster (13.7 / 15.4 vs 13.3 ms)
            for name in ("h_raw", "ya_raw", "yb_raw"):
                b[f"{name}{i}"] = z(B, L + 2 * HALO, cp, dt=F32)
            b[f"xs{i}"] = z(B, L, c0 // 2 ** (i + 1), dt=F32)
        b["audio"] = torch.empty(B, 1, L, dtype=F32, device=dev)
        b["mel_in"] = None

    def _run(self, plan: _Plan, mel: torch.Tensor, time_major: bool, B: int, T: int) -> torch.Tensor:
        h, W, b = self.h, self._w, plan.bufs
        c0 = h.upsample_initial_channel
        L = T
        check(lib().kr_hifi_pack_mel(ops._ptr(mel), ctypes.c_int(int(time_major)), ops._ptr(b["mel_cl"]),
                                     ctypes.c_int(B), ctypes.c_int(T), ctypes.c_int(self.n_mels), ctypes.c_int(HALO),
                                     ctypes.c_int(b["mel_cl"].shape[2]), ops._stream()), "kr_hifi_pack_mel")
        inner = lambda t, n: t[:, HALO:HALO + n]             # noqa: E731
        # conv_pre (k=7, pad 3) -> lrelu 0.1
        ops.conv1d_cl(b["mel_cl"], W["pre.w"], rows=T, row0=HALO - 3, taps=7, dil=1, bias=W["pre.b"],
                      out_act=inner(b["x_act"], T), act_slope=0.1)
        x_act = b["x_act"]
        for i, s in enumerate(h.upsample_rates):
            cout = c0 // 2 ** (i + 1)
            cp = _cpad(cout)
            Lin, L = L, L * s
            # polyphase transposed conv: GEMM row q -> s output rows
            ops.conv1d_cl(x_act, W[f"up{i}.w"], rows=Lin, row0=HALO - 1, taps=3, dil=1, bias=W[f"up{i}.b"],
                          out=inner(b[f"h_raw{i}"], L).view(B, Lin, s * cp),
                          out_act=inner(b[f"h_act{i}"], L).view(B, Lin, s * cp), act_slope=0.1)
            xs = b[f"xs{i}"]
            last_stage = i == self.num_upsamples - 1
            for j, (k, dils) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                p = f"resblocks.{i * self.num_kernels + j}"
                y_raw, y_act = b[f"h_raw{i}"], b[f"h_act{i}"]
                # time-folded views (dilation-1 convs of the narrow stage, see _tap_major_time_folded): [B, rows/2, 2C] over
                # the same memory; f2 = the padded activation buffer, i2 = its inner L rows, x2 = an un-padded [B, L, C] tensor
                hf = ((k - 1) // 2 + 1) // 2
                f2 = lambda t: t.view(B, (L + 2 * HALO) // 2, 2 * cp)                       # noqa: E731
                i2 = lambda t: f2(t)[:, HALO // 2:HALO // 2 + L // 2]                       # noqa: E731
                x2 = lambda t: t.view(B, L // 2, 2 * cout)                                  # noqa: E731
                for d, dil in enumerate(dils):
                    final = d == len(dils) - 1
                    last_rb = j == self.num_kernels - 1
                    n_raw, n_act = (b[f"ya_raw{i}"], b[f"ya_act{i}"]) if d % 2 == 0 else (b[f"yb_raw{i}"], b[f"yb_act{i}"])
                    # destination of this step: the next step's (raw, act) pair, or — last step of the ResBlock — the MRF
                    # accumulator xs (+)= (x + conv) / num_kernels, whose last writer emits lrelu(xs) for the next stage
                    fold1, fold2 = f"{p}.convs1.{d}.w2" in W, f"{p}.convs2.{d}.w2" in W
                    rb = W.get(f"{p}.rb.{d}")
                    if rb is not None:
                        # ONE kernel for c1 -> lrelu -> c2 -> + x; on the narrow stage through the time-folded views
                        fz = rb["folded"]
                        v_in = f2(y_act) if fz else y_act
                        iv = i2 if fz else (lambda t: inner(t, L)[:, :, :cout])
                        xv = x2 if fz else (lambda t: t)
                        kw = dict(resid=iv(y_raw))
                        if not final:
                            kw.update(out=iv(n_raw), out_act=iv(n_act), act_slope=0.1)
                        else:
                            kw.update(resid2=xv(xs) if j > 0 else None, beta=1.0 / self.num_kernels,
                                      out=None if last_rb else xv(xs), out_act=iv(b[f"x_act{i}"]) if last_rb else None,
                                      act_slope=0.01 if last_stage else 0.1)
                        ops.hifi_resblock(v_in, L // 2 if fz else L, HALO // 2 if fz else HALO, rb["w1"], rb["off1"], rb["kh1"],
                                          rb["b1"], rb["w2"], rb["off2"], rb["kh2"], rb["b2"], **kw)
                        if not final:
                            y_raw, y_act = n_raw, n_act
                        continue
                    if fold1:
                        ops.conv1d_cl(f2(y_act), W[f"{p}.convs1.{d}.w2"], rows=L // 2, row0=HALO // 2 - hf, taps=2 * hf + 1,
                                      dil=1, bias=W[f"{p}.convs1.{d}.b2"], out_act=i2(b[f"t_act{i}"]), act_slope=0.1)
                    else:
                        ops.conv1d_cl(y_act, W[f"{p}.convs1.{d}.w"], rows=L, row0=HALO - dil * (k - 1) // 2, taps=k,
                                      dil=dil, bias=W[f"{p}.convs1.{d}.b"], out_act=inner(b[f"t_act{i}"], L)[:, :, :cout],
                                      act_slope=0.1)
                    folded = fold2
                    if not final:
                        if folded:
                            ops.conv1d_cl(f2(b[f"t_act{i}"]), W[f"{p}.convs2.{d}.w2"], rows=L // 2, row0=HALO // 2 - hf,
                                          taps=2 * hf + 1, dil=1, bias=W[f"{p}.convs2.{d}.b2"], resid=i2(y_raw),
                                          out=i2(n_raw), out_act=i2(n_act), act_slope=0.1)
                        else:
                            ops.conv1d_cl(b[f"t_act{i}"], W[f"{p}.convs2.{d}.w"], rows=L, row0=HALO - (k - 1) // 2,
                                          taps=k, dil=1, bias=W[f"{p}.convs2.{d}.b"], resid=inner(y_raw, L)[:, :, :cout],
                                          out=inner(n_raw, L)[:, :, :cout], out_act=inner(n_act, L)[:, :, :cout],
                                          act_slope=0.1)
                        y_raw, y_act = n_raw, n_act
                    else:
                        if folded:
                            ops.conv1d_cl(f2(b[f"t_act{i}"]), W[f"{p}.convs2.{d}.w2"], rows=L // 2, row0=HALO // 2 - hf,
                                          taps=2 * hf + 1, dil=1, bias=W[f"{p}.convs2.{d}.b2"], resid=i2(y_raw),
                                          resid2=x2(xs) if j > 0 else None, beta=1.0 / self.num_kernels,
                                          out=None if last_rb else x2(xs),
                                          out_act=i2(b[f"x_act{i}"]) if last_rb else None,
                                          act_slope=0.01 if last_stage else 0.1)
                        else:
                            ops.conv1d_cl(b[f"t_act{i}"], W[f"{p}.convs2.{d}.w"], rows=L, row0=HALO - (k - 1) // 2,
                                          taps=k, dil=1, bias=W[f"{p}.convs2.{d}.b"], resid=inner(y_raw, L)[:, :, :cout],
                                          resid2=xs if j > 0 else None, beta=1.0 / self.num_kernels,
                                          out=None if last_rb else xs,
                                          out_act=inner(b[f"x_act{i}"], L)[:, :, :cout] if last_rb else None,
                                          act_slope=0.01 if last_stage else 0.1)
            x_act = b[f"x_act{i}"]
        ch = c0 // 2 ** self.num_upsamples
        check(lib().kr_hifi_post_tanh(ops._ptr(x_act), ops._ptr(W["post.w"]), ops._ptr(W["post.b"]),
                                      ops._ptr(b["audio"]), ctypes.c_int(B), ctypes.c_longlong(L), ctypes.c_int(HALO),
                                      ctypes.c_int(ch), ctypes.c_int(x_act.shape[2]), ctypes.c_int(7), ops._stream()),
              "kr_hifi_post_tanh")
        return b["audio"]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: (B, 80, T), (B, T, 80) (auto-detected exactly as the reference :112-117) or (T, 80).
        Returns (B, 1, T * prod(upsample_rates)) fp32 on the device (a buffer owned by the plan: it is
        overwritten by the next forward of the same shape)."""
        if not self._folded:
            self._fold()
        time_major = False
        if x.dim() == 3 and x.size(1) != self.n_mels and x.size(2) == self.n_mels:
            time_major = True
        elif x.dim() == 2:
            x = x.unsqueeze(0)
            time_major = True
        if x.dim() != 3:
            raise ValueError(f"unsupported mel shape {tuple(x.shape)}")
        B = x.size(0)
        T = x.size(1) if time_major else x.size(2)
        key = (B, T, time_major)
        plan = self._plans.get(key)
        if plan is None:
            plan = _Plan()
            self._alloc(plan, B, T)
            plan.bufs["mel_in"] = torch.empty(x.shape, dtype=F32, device=self.device)
            self._plans[key] = plan
        mel_in = plan.bufs["mel_in"]
        mel_in.copy_(x, non_blocking=True)                    # H2D (or D2D) into the static input buffer
        if not self.use_graphs or not plan.warm:
            n0 = launch_count()
            out = self._run(plan, mel_in, time_major, B, T)
            plan.launches = launch_count() - n0
            plan.warm = True
        else:
            if plan.graph is None:
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run(plan, mel_in, time_major, B, T)
                plan.graph = g
            plan.graph.replay()
            out = plan.bufs["audio"]
        self.launches_last_forward = plan.launches
        return out

    __call__ = forward


def load_hifigan_model(model_path, config_path=None, device: str = "cuda") -> HiFiGANGenerator:
    """reference hifigan_vocoder.py:190-271: accepts {'generator': sd} or a raw state dict."""
    config = HiFiGANConfig.load_config(config_path) if config_path and Path(config_path).exists() \
        else HiFiGANConfig.get_default_config()
    gen = HiFiGANGenerator(config, device=device)
    sd = torch.load(model_path, map_location="cpu", weights_only=True)
    gen.load_state_dict(sd["generator"] if "generator" in sd else sd)
    return gen


class VocoderManager:
    """HiFi-GAN side of the reference's ``VocoderManager`` (inference/vocoder_manager.py:22-206): same constructor
    arguments, ``mel_to_audio`` with the reference's layout rules and its 1-D / 2-D CPU result.  Griffin-Lim (the
    reference's CPU fallback when no checkpoint is available) is deliberately NOT provided: this path has no CPU
    fallback, and model downloads need a network.  ``vocoder`` may be injected (tests, pre-loaded generators)."""

    def __init__(self, vocoder_type: str = "hifigan", vocoder_path: Optional[str] = None, device: str = "cuda",
                 config_path: Optional[str] = None, vocoder=None):
        self.vocoder_type = vocoder_type.lower()
        self.device = device
        if self.vocoder_type != "hifigan":
            raise ValueError(f"Unsupported vocoder type on the B200 path: {vocoder_type} (HiFi-GAN only, no Griffin-Lim fallback)")
        if vocoder is not None:
            self.vocoder = vocoder
        elif vocoder_path is not None:
            self.vocoder = load_hifigan_model(vocoder_path, config_path, device=device)
        else:
            raise RuntimeError("VocoderManager needs a local HiFi-GAN checkpoint (vocoder_path): downloads are not supported")

    def mel_to_audio(self, mel_spec: torch.Tensor) -> torch.Tensor:
        """vocoder_manager.py:154-206: (n_mels, T) -> add a batch dim; a 3-D input whose first dim is not 1 is taken as
        (batch, T, n_mels) and transposed (a batch of ONE is passed through unchanged and left to the generator's own
        layout detection, hifigan_vocoder.py:112-117 — the reference's behaviour); the result loses its batch and
        channel dims when they are 1 and is returned on the CPU."""
        with torch.no_grad():
            mel_spec = mel_spec.to(self.device)
            if mel_spec.dim() == 2:
                mel_spec = mel_spec.unsqueeze(0)
            elif mel_spec.dim() == 3 and mel_spec.shape[0] != 1:
                mel_spec = mel_spec.transpose(1, 2)
            audio = self.vocoder(mel_spec)
            if audio.dim() == 3:
                audio = audio.squeeze(0)
            if audio.dim() == 2:
                audio = audio.squeeze(0)
        return audio.cpu()
