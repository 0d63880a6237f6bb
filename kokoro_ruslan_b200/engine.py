"""Acoustic-model training forward / backward as a sequence of libkokoro_b200 kernel launches.

Mirrors the data flow of the reference ``KokoroModel.forward_training``
(src/kokoro/model/model.py:565-673 and SURVEY.md §9 S1-S9): embedding -> 6 pre-norm encoder
blocks -> variance adaptor (duration predictor, detached length regulation, pitch/energy
predictors and embeddings) -> teacher-forced 6-block decoder -> mel / stop heads; then the fused
losses and a hand-scheduled backward over the four disjoint sub-graphs the reference's two
``detach()`` cuts create.  Host code only allocates tensors and orders launches; all arithmetic is
in the CUDA library (tcgen05 GEMM / flash attention + HBM kernels).  Dropout and stochastic depth
(``DropoutConfig``) are fused into those kernels through a counter-based RNG: masks are functions of
(seed, step, site, element) and are regenerated, never stored (csrc/kr_common.cuh).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .params import ModelConfig, ParamStore

F32, BF16 = torch.float32, torch.bfloat16


@dataclass
class LossConfig:
    """Loss weights / criteria constants (reference training/config.py:133-141, trainer.py:410-444)."""
    w_dur: float = 0.35
    w_stop: float = 0.010
    w_pitch: float = 1.0
    w_energy: float = 1.0
    stop_pos_weight: float = 17.0
    huber_delta_var: float = 0.05


@dataclass
class DropoutConfig:
    """Drop probabilities of the reference model (KokoroModel ctor, model/model.py:35-47,78; the
    trainer's values are training/config.py:108-121,195).  All zeros = the deterministic parity
    configuration."""
    encoder: float = 0.0            # encoder attention / FFN / residual dropouts AND both positional-encoding dropouts
    decoder: float = 0.0            # decoder attention / FFN / residual dropouts
    decoder_input: float = 0.0      # F.dropout on the projected, shifted mel (model.py:525)
    variance: float = 0.0           # VariancePredictor conv stack (variance_predictor.py:106)
    stochastic_depth: float = 0.0   # max drop-path rate, linear over the layers (model.py:99-107)
    seed: int = 0

    @staticmethod
    def reference_training() -> "DropoutConfig":
        """TrainingConfig defaults: encoder 0.15, decoder 0.20, decoder input 0.15, variance 0.1, stochastic depth 0.1."""
        return DropoutConfig(encoder=0.15, decoder=0.20, decoder_input=0.15, variance=0.1, stochastic_depth=0.1)

    def any(self) -> bool:
        return any(v > 0 for v in (self.encoder, self.decoder, self.decoder_input, self.variance, self.stochastic_depth))


class PadGeom:
    """Row tables of the predictor 'padded layout' for B sequences of length L split into
    independent chunks (reference model/variance_predictor.py:77-87): every chunk is surrounded by
    one zero row on each side (= the conv's zero padding)."""

    def __init__(self, B: int, L: int, chunk: int, device):
        self.B, self.L, self.chunk = B, L, chunk
        row_group: List[int] = []
        tok_of_row: List[int] = []
        row_of_tok = [0] * (B * L)
        group_rows: List[int] = []
        for b in range(B):
            for s in range(0, L, chunk):
                n = min(chunk, L - s)
                gid = len(group_rows)
                group_rows.append(n)
                row_group.append(-1)
                tok_of_row.append(-1)
                for t in range(s, s + n):
                    row_of_tok[b * L + t] = len(row_group)
                    row_group.append(gid)
                    tok_of_row.append(b * L + t)
                row_group.append(-1)
                tok_of_row.append(-1)
        self.R = len(row_group)
        self.G = len(group_rows)
        i32 = dict(dtype=torch.int32, device=device)
        self.row_group = torch.tensor(row_group, **i32)
        self.tok_of_row = torch.tensor(tok_of_row, **i32)
        self.row_of_tok = torch.tensor(row_of_tok, **i32)
        self.group_rows = torch.tensor(group_rows, **i32)


# Weight-gradient GEMM policy, fixed by sweeps on B200 at the bench shape (tools/gpu_ab.sh history in DESIGN.md §7):
# 128 x 256 output tiles (one N = 256 MMA per K step reads fewer operand bytes per flop than two N = 128 ones), split-K
# until a GEMM has about _WGRAD_TARGET CTAs, never fewer than _WGRAD_MIN_KB 64-row K blocks per split.
_WGRAD_TARGET = 112
_WGRAD_MIN_KB = 12
_WGRAD_BLOCK_N = 256


def _auto_splits(tiles: int, k_blocks: int, target: int = 0) -> int:
    """Split-K factor of a weight-gradient GEMM.  Every split pays a full fp32-atomic epilogue of its output tile
    and the weight gradients run on side streams next to the critical chain, so FEWER, longer CTAs win: measured
    on B200 (tools/wgrad_ab.sh) 296 CTAs / no minimum = 7.36 ms per step, 96 CTAs with >= 12 K blocks (768 rows)
    per split = 7.07-7.16 ms.  Round 2, 128 x 256 tiles, step in ms at target 96 / 112 / 128 / 148 / 192 CTAs:
    6.19 / 6.14 / 6.16 / 6.21 / 6.20 (128 x 128 tiles at 96 / 128: 6.19 / 6.20); a finer sweep on the final tree: 104 / 112 /
    120 CTAs = 6.140 / 6.144 - 6.150 / 6.150, minimum K blocks per split 8 / 12 / 16 = 6.186 / 6.144 / 6.158."""
    target = target or _WGRAD_TARGET
    s = max(1, min(k_blocks, (target + tiles - 1) // tiles))
    return max(1, min(s, k_blocks // max(1, _WGRAD_MIN_KB)))


class AcousticEngine:
    def __init__(self, cfg: ModelConfig, device=None, with_ema: bool = True,
                 loss_cfg: Optional[LossConfig] = None, multi_stream: bool = True,
                 dropout: Optional[DropoutConfig] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("AcousticEngine needs a CUDA device: the hot path has no CPU fallback")
        self.cfg = cfg
        self.device = torch.device(device if device is not None else "cuda")
        self.loss_cfg = loss_cfg or LossConfig()
        self.store = ParamStore(cfg, self.device, with_ema=with_ema)
        self.H = cfg.n_heads
        self.D = cfg.hidden_dim
        assert self.D == self.H * 64, "head_dim must be 64 (tcgen05 attention tiles)"
        self._geoms: Dict[Tuple[int, int], PadGeom] = {}
        # Independent sub-graphs run on side streams (fork/join from the caller's stream, so the
        # whole step is still one capturable DAG): weight-gradient GEMMs + bias column sums (off
        # the critical path), the variance predictors, and the encoder backward, which is disjoint
        # from the decoder backward because the length regulator detaches (utils/lengths.py:30).
        import os
        if os.environ.get("KR_STREAMS", "1") == "0":
            multi_stream = False
        self.multi_stream = multi_stream
        self._side = {k: torch.cuda.Stream(device=self.device) for k in ("w0", "w1", "vp", "enc", "kv", "d0", "comm", "z", "l")} if multi_stream else {}
        # data parallel: callable(split_layer) that all-reduces early_grad_ranges(split_layer); backward_parts() runs it on
        # the "comm" side stream once those ranges are final, underneath the rest of the backward (TrainStep installs it)
        self.early_reduce_hook = None
        self._w_rr = 0
        self._forked: List[torch.cuda.Stream] = []
        # every tensor of a step stays referenced until the next step starts: memory is never
        # recycled across streams inside a step (the caching allocator is per-stream ordered).
        self._live: List[torch.Tensor] = []
        # SpecAugment span table (device int32 [B, n_time + n_feat, 2]) or None; see set_spec_augment()
        self.spec_spans: Optional[torch.Tensor] = None
        self.spec_n_time = self.spec_n_feat = 0
        self.training = True
        self._init_dropout(dropout)

    # ------------------------------------------------------------------------------------------
    # dropout / stochastic depth
    # ------------------------------------------------------------------------------------------
    def _init_dropout(self, dropout: Optional[DropoutConfig]) -> None:
        """Site table: every dropout of the reference forward gets a fixed integer id (the RNG stream
        selector); residual branches additionally get a row of the per-step stochastic-depth table."""
        cfg = self.cfg
        self.dropout = dropout if (dropout is not None and dropout.any()) else None
        names: List[str] = ["enc.pe", "dec.in", "dec.pe"]
        path: List[Tuple[str, float]] = []
        sd = self.dropout.stochastic_depth if self.dropout is not None else 0.0
        for i in range(cfg.n_encoder_layers):
            rate = (i / max(cfg.n_encoder_layers - 1, 1)) * sd
            for sub in ("attn", "ffn"):
                names += [f"enc.{i}.{sub}.p", f"enc.{i}.{sub}.u", f"enc.{i}.{sub}.out", f"enc.{i}.{sub}.out2",
                          f"enc.{i}.{sub}.path"]
                path.append((f"enc.{i}.{sub}.path", rate))
        for i in range(cfg.n_decoder_layers):
            rate = (i / max(cfg.n_decoder_layers - 1, 1)) * sd
            for sub in ("self", "cross", "ffn"):
                names += [f"dec.{i}.{sub}.p", f"dec.{i}.{sub}.u", f"dec.{i}.{sub}.out", f"dec.{i}.{sub}.out2",
                          f"dec.{i}.{sub}.path"]
                path.append((f"dec.{i}.{sub}.path", rate))
        for vp in ("duration", "pitch", "energy"):
            names += [f"vp.{vp}.0", f"vp.{vp}.1"]
        self.drop_sites: Dict[str, int] = {n: i + 1 for i, n in enumerate(names)}
        self._path_rows: Dict[str, int] = {n: r for r, (n, _) in enumerate(path)}
        self._path_rates = [r for _, r in path]
        self._path_table: Optional[torch.Tensor] = None
        if self.dropout is None:
            self.drop_state = None
            return
        self.drop_state = torch.tensor([int(self.dropout.seed), 0], dtype=torch.int64, device=self.device)
        self._path_site_dev = torch.tensor([self.drop_sites[n] for n, _ in path], dtype=torch.int32, device=self.device)
        self._path_p_dev = torch.tensor(self._path_rates, dtype=torch.float32, device=self.device)

    def set_dropout(self, dropout: Optional[DropoutConfig]) -> None:
        self._init_dropout(dropout)

    @property
    def _drop_on(self) -> bool:
        return self.training and self.dropout is not None

    def _ds(self, site: Optional[str], p: float = 0.0, site_b: Optional[str] = None, p_b: float = 0.0,
            path: Optional[str] = None, rows_per_sample: int = 1, byte_lanes: bool = False):
        """Drop spec of one site (None when dropout is off or the site is a no-op)."""
        if not self._drop_on:
            return None
        row = None
        if path is not None and self._path_rates[self._path_rows[path]] > 0:
            row = self._path_table[self._path_rows[path]]
        return ops.make_drop_spec(self.drop_state, self.drop_sites[site] if site else 0, p,
                                  self.drop_sites[site_b] if site_b else 0, p_b, row, rows_per_sample, byte_lanes)

    def _branch_specs(self, tag: str, p: float, S: int, ffn: bool) -> dict:
        """Specs of one residual branch `tag` ('enc.3.attn', 'dec.0.ffn', ...) with rows_per_sample = S:
        'p' attention probabilities, 'u' FFN inner dropout, 'out' branch output (the block's dropout, for an
        FFN preceded by the FFN's own output dropout, and the stochastic-depth factor)."""
        if not self._drop_on:
            return {"p": None, "u": None, "out": None}
        out = self._ds(f"{tag}.out", p, f"{tag}.out2" if ffn else None, p if ffn else 0.0, path=f"{tag}.path",
                       rows_per_sample=S)
        return {"p": None if ffn else self._ds(f"{tag}.p", p, byte_lanes=True), "u": self._ds(f"{tag}.u", p) if ffn else None,
                "out": out}

    # ------------------------------------------------------------------------------------------
    def _geom(self, B: int, L: int) -> PadGeom:
        key = (B, L)
        if key not in self._geoms:
            self._geoms[key] = PadGeom(B, L, self.cfg.vp_chunk, self.device)
        return self._geoms[key]

    def _empty(self, *shape, dtype=F32):
        t = torch.empty(*shape, dtype=dtype, device=self.device)
        self._live.append(t)
        return t

    def _zeros(self, *shape, dtype=F32):
        t = ops.zero_(torch.empty(*shape, dtype=dtype, device=self.device))
        self._live.append(t)
        return t

    # ------------------------------------------------------------------------------------------
    # stream fork / join
    # ------------------------------------------------------------------------------------------
    class _On:
        """`with eng._on("vp"):` runs the body on a side stream that first waits for everything
        enqueued so far on the current stream; _join_all() makes the current stream wait for it."""

        def __init__(self, eng, name):
            self.eng, self.name, self.ctx = eng, name, None

        def __enter__(self):
            eng = self.eng
            if not eng.multi_stream:
                return self
            side = eng._side[self.name]
            side.wait_stream(torch.cuda.current_stream(eng.device))
            if side not in eng._forked:
                eng._forked.append(side)
            self.ctx = torch.cuda.stream(side)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.ctx is not None:
                self.ctx.__exit__(*exc)
            return False

    def _on(self, name: str):
        return AcousticEngine._On(self, name)

    def _on_w(self):
        """Round-robin over the two weight-gradient streams.  The encoder branch (enqueued first) shares them with the decoder
        chain, so a decoder weight gradient queued behind an encoder one waits until the encoder backward has reached that
        point: in the CUPTI timeline (tools/step_timeline.py) most decoder weight gradients start only when the encoder
        chain has finished (4.85 of 6.0 ms) and 0.26 ms of them trail the decoder chain.  That accidental throttle is the
        best schedule measured: a separate stream for the encoder branch's weight gradients (they then compete with the
        decoder chain for the SMs from the start: the chain ends 0.3 ms later) gave 6.33 vs 6.28 ms per step, deferring all
        decoder weight gradients behind the chain 6.48 ms, a high-priority capture stream for the chain 6.56 ms — the
        backward pass is bound by total SM time, not by the length of any one chain."""
        self._w_rr ^= 1
        return self._on("w1" if self._w_rr else "w0")

    def _join(self, name: str):
        """The current stream waits for side stream `name` only."""
        if not self.multi_stream:
            return
        side = self._side[name]
        torch.cuda.current_stream(self.device).wait_stream(side)
        if side in self._forked:
            self._forked.remove(side)

    def _join_all(self):
        cur = torch.cuda.current_stream(self.device)
        for side in self._forked:
            cur.wait_stream(side)
        self._forked = []

    # ------------------------------------------------------------------------------------------
    # linear helpers
    # ------------------------------------------------------------------------------------------
    def _wgrad(self, dy: torch.Tensor, x: torch.Tensor, gw: torch.Tensor, gb: Optional[torch.Tensor] = None):
        """gw[N_out, K_in] += dy[tok, N_out]^T x[tok, K_in] (split-K, fp32 atomics) and, when gb is
        given, gb[N_out] += column sums of dy (bias gradient) — both on a weight-gradient stream."""
        n_out, k_in = gw.shape
        bn = _WGRAD_BLOCK_N if k_in % _WGRAD_BLOCK_N == 0 else 0          # 0: the library picks the tile width
        tiles = ((n_out + 127) // 128) * ((k_in + (bn or 128) - 1) // (bn or 128))
        kb = (dy.shape[0] + 63) // 64
        with self._on_w():
            ops.gemm(dy, x, gw, a_mn_major=True, b_mn_major=True, accumulate=True,
                     splits=_auto_splits(tiles, kb), block_n=bn)
            if gb is not None:
                ops.colsum_bf16(dy, gb)

    # ------------------------------------------------------------------------------------------
    # attention sub-layer
    # ------------------------------------------------------------------------------------------
    def _cross_kv(self, pre: str, mem: torch.Tensor, Nk: int, Sk: int):
        """K/V projection + per-head RMSNorm of the cross-attention memory (no RoPE on cross-attention)."""
        st, D, H = self.store, self.D, self.H
        raw_kv = self._empty(Nk, 2 * D, dtype=BF16)
        ops.gemm(mem, st.span(st.shadow, pre + "w_k.weight", 2 * D, D), raw_kv)
        nkv = self._empty(Nk, 2 * D, dtype=BF16)
        ops.qkv_prep_fwd([raw_kv[:, :D], raw_kv[:, D:]], [nkv[:, :D], nkv[:, D:]],
                         [st.p(pre + "k_norm.weight"), st.p(pre + "v_norm.weight")], 0, st.rope_cos, st.rope_sin, Nk, Sk, H)
        return raw_kv, nkv

    def _attn_fwd(self, pre: str, x: torch.Tensor, B: int, S: int, norm: str, causal: bool,
                  key_mask: Optional[torch.Tensor], mem: Optional[torch.Tensor], Sk: int, sv: dict, kv_pre=None,
                  drop: Optional[dict] = None, pre_ln=None, next_ln=None):
        """pre_ln = (h, mean, rstd): this sub-layer's LayerNorm was already produced by the tail kernel of the sub-layer
        before it.  next_ln = (norm name, "bf16" | "f32"): the tail of THIS sub-layer (dropout + residual) also produces the
        LayerNorm that follows — returns (out, (h, mean, rstd)) instead of out."""
        st, D, H = self.store, self.D, self.H
        N = B * S
        cross = mem is not None
        if pre_ln is None:
            h = self._empty(N, D, dtype=BF16)
            mean, rstd = self._empty(N), self._empty(N)
            ops.layernorm_fwd(x, st.p(norm + "weight"), st.p(norm + "bias"), h, None, mean, rstd)
        else:
            h, mean, rstd = pre_ln
        gq, gk, gv = st.p(pre + "q_norm.weight"), st.p(pre + "k_norm.weight"), st.p(pre + "v_norm.weight")
        if not cross:
            raw = self._empty(N, 3 * D, dtype=BF16)
            ops.gemm(h, st.span(st.shadow, pre + "w_q.weight", 3 * D, D), raw)
            nrm = self._empty(N, 3 * D, dtype=BF16)
            parts_in = [raw[:, :D], raw[:, D:2 * D], raw[:, 2 * D:]]
            parts_out = [nrm[:, :D], nrm[:, D:2 * D], nrm[:, 2 * D:]]
            ops.qkv_prep_fwd(parts_in, parts_out, [gq, gk, gv], 0b011, st.rope_cos, st.rope_sin, N, S, H)
            q, k, v = (t.view(B, S, H, 64) for t in parts_out)
            sv.update(raw=raw, nrm=nrm)
        else:
            Nk = B * Sk
            raw_q = self._empty(N, D, dtype=BF16)
            ops.gemm(h, st.w(pre + "w_q.weight"), raw_q)
            nq = self._empty(N, D, dtype=BF16)
            ops.qkv_prep_fwd([raw_q], [nq], [gq], 0, st.rope_cos, st.rope_sin, N, S, H)
            if kv_pre is not None:            # K/V of the memory were projected ahead of time on the "kv" stream
                raw_kv, nkv, ev = kv_pre
                if ev is not None:
                    torch.cuda.current_stream(self.device).wait_event(ev)
            else:
                raw_kv, nkv = self._cross_kv(pre, mem, Nk, Sk)
            q = nq.view(B, S, H, 64)
            k, v = nkv[:, :D].view(B, Sk, H, 64), nkv[:, D:].view(B, Sk, H, 64)
            sv.update(raw_q=raw_q, raw_kv=raw_kv, nq=nq, nkv=nkv)
        o = self._empty(N, D, dtype=BF16)
        lse = self._empty(B, H, S)
        drop = drop or {"p": None, "out": None}
        ops.attn_fwd(q, k, v, o.view(B, S, H, 64), lse, key_mask, causal, 1.0 / 8.0, drop=drop["p"])
        out = self._empty(N, D)
        nxt = None
        if next_ln is None:
            ops.gemm(o, st.w(pre + "w_o.weight"), out, bias=st.p(pre + "w_o.bias"), resid=x, drop=drop["out"])
        else:
            # plain out-projection (its dropout / residual epilogue costs 8 us of an exposed single-tile epilogue), then ONE
            # row-wise kernel: dropout + residual + the LayerNorm of the sub-layer that follows
            yo = self._empty(N, D)
            ops.gemm(o, st.w(pre + "w_o.weight"), yo, bias=st.p(pre + "w_o.bias"))
            nxt = self._ln_out(N, next_ln)
            ops.resid_drop_ln_fwd(yo, x, out, drop["out"], st.p(next_ln[0] + "weight"), st.p(next_ln[0] + "bias"), *nxt)
            nxt = (nxt[0] if nxt[0] is not None else nxt[1], nxt[2], nxt[3])
        sv.update(x=x, h=h, mean=mean, rstd=rstd, o=o, lse=lse, q=q, k=k, v=v, drop=drop)
        return out if next_ln is None else (out, nxt)

    def _ln_out(self, N: int, next_ln):
        """(h_bf16 | None, h_f32 | None, mean, rstd) buffers of a fused following LayerNorm."""
        bf = next_ln[1] == "bf16"
        return (self._empty(N, self.D, dtype=BF16) if bf else None, None if bf else self._empty(N, self.D),
                self._empty(N), self._empty(N))

    def _attn_bwd(self, pre: str, dout: torch.Tensor, dout_bf: torch.Tensor, B: int, S: int, norm: str,
                  causal: bool, key_mask, mem, Sk: int, sv: dict, dmem: Optional[torch.Tensor],
                  dmem_first: bool, next_drop=None, next_dbias=None, bias_done: bool = False):
        """dout_bf must already carry this branch's output dropout (sv['drop']['out']): it is produced by the
        layernorm_bwd of the sub-layer that follows in the forward.  next_drop = the output-dropout spec of the
        branch that PRECEDES this one in the forward, applied to the returned dx_bf."""
        st, D, H = self.store, self.D, self.H
        N = B * S
        cross = mem is not None
        # bias_done: the layernorm_bwd that produced dout_bf already accumulated its column sums into w_o.bias
        self._wgrad(dout_bf, sv["o"], st.g(pre + "w_o.weight"), None if bias_done else st.g(pre + "w_o.bias"))
        d_o = self._empty(N, D, dtype=BF16)
        ops.gemm(dout_bf, st.w(pre + "w_o.weight"), d_o, b_mn_major=True)
        dq = self._empty(N, D)            # zeroed by kr_attn_bwd's prep kernel
        Nk = B * Sk
        dkv = self._empty(Nk, 2 * D, dtype=BF16)
        delta = self._empty(B, H, S)
        ops.attn_bwd(sv["q"], sv["k"], sv["v"], sv["o"].view(B, S, H, 64), d_o.view(B, S, H, 64), sv["lse"],
                     delta, dq.view(B, S, H, 64), dkv[:, :D].view(B, Sk, H, 64), dkv[:, D:].view(B, Sk, H, 64),
                     key_mask, causal, 1.0 / 8.0, drop=sv["drop"]["p"])
        gq, gk, gv = st.p(pre + "q_norm.weight"), st.p(pre + "k_norm.weight"), st.p(pre + "v_norm.weight")
        dgq, dgk, dgv = st.g(pre + "q_norm.weight"), st.g(pre + "k_norm.weight"), st.g(pre + "v_norm.weight")
        dh = self._empty(N, D)
        if not cross:
            raw = sv["raw"]
            draw = self._empty(N, 3 * D, dtype=BF16)
            ops.qkv_prep_bwd([raw[:, :D], raw[:, D:2 * D], raw[:, 2 * D:]], [dq, dkv[:, :D], dkv[:, D:]],
                             [draw[:, :D], draw[:, D:2 * D], draw[:, 2 * D:]], [gq, gk, gv], [dgq, dgk, dgv],
                             0b011, st.rope_cos, st.rope_sin, N, S, H)
            self._wgrad(draw, sv["h"], st.span(st.grads, pre + "w_q.weight", 3 * D, D))
            ops.gemm(draw, st.span(st.shadow, pre + "w_q.weight", 3 * D, D), dh, b_mn_major=True)
        else:
            # K/V side of the cross-attention: only feeds weight gradients and the memory gradient, which is
            # consumed after the last layer -> off the critical path, serialised on the "kv" stream
            with self._on("kv"):
                raw_kv = sv["raw_kv"]
                dkv_raw = self._empty(Nk, 2 * D, dtype=BF16)
                ops.qkv_prep_bwd([raw_kv[:, :D], raw_kv[:, D:]], [dkv[:, :D], dkv[:, D:]],
                                 [dkv_raw[:, :D], dkv_raw[:, D:]], [gk, gv], [dgk, dgv], 0, st.rope_cos,
                                 st.rope_sin, Nk, Sk, H)
                self._wgrad(dkv_raw, mem, st.span(st.grads, pre + "w_k.weight", 2 * D, D))
                ops.gemm(dkv_raw, st.span(st.shadow, pre + "w_k.weight", 2 * D, D), dmem, b_mn_major=True,
                         resid=None if dmem_first else dmem)
            dq_raw = self._empty(N, D, dtype=BF16)
            ops.qkv_prep_bwd([sv["raw_q"]], [dq], [dq_raw], [gq], [dgq], 0, st.rope_cos, st.rope_sin, N, S, H)
            self._wgrad(dq_raw, sv["h"], st.g(pre + "w_q.weight"))
            ops.gemm(dq_raw, st.w(pre + "w_q.weight"), dh, b_mn_major=True)
        dx = self._empty(N, D)
        dx_bf = self._empty(N, D, dtype=BF16)
        ops.layernorm_bwd(dh, sv["x"], sv["mean"], sv["rstd"], st.p(norm + "weight"), dout, dx, dx_bf,
                          st.g(norm + "weight"), st.g(norm + "bias"), drop_bf16=next_drop, dcol_bf16=next_dbias)
        return dx, dx_bf

    # ------------------------------------------------------------------------------------------
    # GLU feed-forward sub-layer
    # ------------------------------------------------------------------------------------------
    def _ffn_fwd(self, pre: str, x: torch.Tensor, norm: str, ff: int, sv: dict, drop: Optional[dict] = None, pre_ln=None,
                 next_ln=None):
        """pre_ln / next_ln: as in _attn_fwd."""
        st, D = self.store, self.D
        N = x.shape[0]
        if pre_ln is None:
            h = self._empty(N, D, dtype=BF16)
            mean, rstd = self._empty(N), self._empty(N)
            ops.layernorm_fwd(x, st.p(norm + "weight"), st.p(norm + "bias"), h, None, mean, rstd)
        else:
            h, mean, rstd = pre_ln
        hff = self._empty(N, 2 * ff, dtype=BF16)
        ops.gemm(h, st.w(pre + "linear1.weight"), hff, bias=st.p(pre + "linear1.bias"))
        u = self._empty(N, ff, dtype=BF16)
        drop = drop or {"u": None, "out": None}
        ops.glu_fwd(hff, u, drop=drop["u"])
        y = self._empty(N, D)
        ops.gemm(u, st.w(pre + "linear2.weight"), y, bias=st.p(pre + "linear2.bias"))
        out = self._empty(N, D)
        nxt = None
        if next_ln is None:
            ops.rmsnorm_resid_fwd(y, st.p(pre + "output_norm.weight"), x, out, drop=drop["out"])
        else:
            nxt = self._ln_out(N, next_ln)
            ops.rmsnorm_resid_ln_fwd(y, st.p(pre + "output_norm.weight"), x, out, drop["out"], st.p(next_ln[0] + "weight"),
                                     st.p(next_ln[0] + "bias"), *nxt)
            nxt = (nxt[0] if nxt[0] is not None else nxt[1], nxt[2], nxt[3])
        sv.update(x=x, h=h, mean=mean, rstd=rstd, hff=hff, u=u, y=y, drop=drop)
        return out if next_ln is None else (out, nxt)

    def _ffn_bwd(self, pre: str, dout: torch.Tensor, norm: str, ff: int, sv: dict, next_drop=None, next_dbias=None):
        st, D = self.store, self.D
        N = dout.shape[0]
        dy = self._empty(N, D, dtype=BF16)
        ops.rmsnorm_resid_bwd(dout, sv["y"], st.p(pre + "output_norm.weight"), dy, st.g(pre + "output_norm.weight"),
                              drop=sv["drop"]["out"], dcol=st.g(pre + "linear2.bias"))
        self._wgrad(dy, sv["u"], st.g(pre + "linear2.weight"), None)
        du = self._empty(N, ff, dtype=BF16)
        ops.gemm(dy, st.w(pre + "linear2.weight"), du, b_mn_major=True)
        dhff = self._empty(N, 2 * ff, dtype=BF16)
        ops.glu_bwd(du, sv["hff"], dhff, drop=sv["drop"]["u"])
        self._wgrad(dhff, sv["h"], st.g(pre + "linear1.weight"), st.g(pre + "linear1.bias"))
        dh = self._empty(N, D)
        ops.gemm(dhff, st.w(pre + "linear1.weight"), dh, b_mn_major=True)
        dx = self._empty(N, D)
        dx_bf = self._empty(N, D, dtype=BF16)
        ops.layernorm_bwd(dh, sv["x"], sv["mean"], sv["rstd"], st.p(norm + "weight"), dout, dx, dx_bf,
                          st.g(norm + "weight"), st.g(norm + "bias"), drop_bf16=next_drop, dcol_bf16=next_dbias)
        return dx, dx_bf

    # ------------------------------------------------------------------------------------------
    # variance predictor (two k=3 convs as overlapping-row GEMMs + GroupNorm/ReLU + linear head)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _conv_view(buf_guarded: torch.Tensor, R: int, C: int) -> torch.Tensor:
        """[R, 3C] view whose row r spans guarded rows r..r+2 (= padded rows r-1, r, r+1)."""
        return torch.as_strided(buf_guarded, (R, 3 * C), (C, 1))

    def _vp_fwd(self, pre: str, xg: torch.Tensor, geom: PadGeom, mask: Optional[torch.Tensor], sv: dict,
                tag: str = ""):
        st, Fv = self.store, self.cfg.variance_filter_size
        pv = self.dropout.variance if self._drop_on else 0.0
        d1, d2 = self._ds(f"vp.{tag}.0", pv), self._ds(f"vp.{tag}.1", pv)
        R, Cin = geom.R, xg.shape[1]
        a1 = self._conv_view(xg, R, Cin)
        c1 = self._empty(R, Fv)
        ops.gemm(a1, st.w(pre + "conv_layers.0.weight"), c1, bias=st.p(pre + "conv_layers.0.bias"))
        h1g = self._zeros(R + 2, Fv, dtype=BF16)
        stats1 = self._empty(geom.G, 2, dtype=torch.float64)
        ops.gn_fwd(c1, geom.row_group, geom.group_rows, stats1, st.p(pre + "norms.0.weight"),
                   st.p(pre + "norms.0.bias"), h1g[1:R + 1], drop=d1)
        a2 = self._conv_view(h1g, R, Fv)
        c2 = self._empty(R, Fv)
        ops.gemm(a2, st.w(pre + "conv_layers.1.weight"), c2, bias=st.p(pre + "conv_layers.1.bias"))
        h2 = self._empty(R, Fv, dtype=BF16)
        stats2 = self._empty(geom.G, 2, dtype=torch.float64)
        ops.gn_fwd(c2, geom.row_group, geom.group_rows, stats2, st.p(pre + "norms.1.weight"),
                   st.p(pre + "norms.1.bias"), h2, drop=d2)
        out = self._empty(geom.B, geom.L)
        ops.vp_head_fwd(h2, geom.row_of_tok, st.p(pre + "linear.weight"), st.p(pre + "linear.bias"), mask, out,
                        geom.L, self.cfg.vp_chunk)
        sv.update(xg=xg, c1=c1, h1g=h1g, stats1=stats1, c2=c2, h2=h2, stats2=stats2, mask=mask, d1=d1, d2=d2)
        return out

    def _vp_bwd(self, pre: str, dout: torch.Tensor, geom: PadGeom, sv: dict, need_dx: bool):
        st, Fv = self.store, self.cfg.variance_filter_size
        R, Cin = geom.R, sv["xg"].shape[1]
        dh2 = self._empty(R, Fv, dtype=BF16)
        ops.vp_head_bwd(dout, sv["h2"], geom.tok_of_row, st.p(pre + "linear.weight"), sv["mask"], dh2,
                        st.g(pre + "linear.weight"), st.g(pre + "linear.bias"), geom.L, self.cfg.vp_chunk)
        gsum = self._empty(geom.G, 2, dtype=torch.float64)
        dc2g = self._zeros(R + 2, Fv, dtype=BF16)
        ops.gn_bwd(dh2, sv["c2"], geom.row_group, geom.group_rows, sv["stats2"], gsum, st.p(pre + "norms.1.weight"),
                   st.p(pre + "norms.1.bias"), dc2g[1:R + 1], st.g(pre + "norms.1.weight"), st.g(pre + "norms.1.bias"),
                   drop=sv["d2"])
        self._wgrad(dc2g[1:R + 1], self._conv_view(sv["h1g"], R, Fv), st.g(pre + "conv_layers.1.weight"),
                    st.g(pre + "conv_layers.1.bias"))
        dh1 = self._empty(R, Fv, dtype=BF16)
        ops.gemm(self._conv_view(dc2g, R, Fv), st.conv_dgrad[pre + "conv_layers.1.weight"], dh1)
        dc1g = self._zeros(R + 2, Fv, dtype=BF16)
        ops.gn_bwd(dh1, sv["c1"], geom.row_group, geom.group_rows, sv["stats1"], gsum, st.p(pre + "norms.0.weight"),
                   st.p(pre + "norms.0.bias"), dc1g[1:R + 1], st.g(pre + "norms.0.weight"), st.g(pre + "norms.0.bias"),
                   drop=sv["d1"])
        self._wgrad(dc1g[1:R + 1], self._conv_view(sv["xg"], R, Cin), st.g(pre + "conv_layers.0.weight"),
                    st.g(pre + "conv_layers.0.bias"))
        if not need_dx:
            return None
        dxp = self._empty(R, Cin)
        ops.gemm(self._conv_view(dc1g, R, Fv), st.conv_dgrad[pre + "conv_layers.0.weight"], dxp)
        dx = self._empty(geom.B * geom.L, Cin)
        ops.gather_rows(dxp, geom.row_of_tok, dx)
        return dx

    def _decoder_head(self, ctx: dict, mel_specs, B: int, T: int, dcfg, p_enc: float, p_dec: float):
        """Decoder input (shift-right, mel_projection_in, dropouts, PE) and the self-attention sub-layer of decoder
        layer 0: the only part of the decoder that does not depend on the encoder, so forward() runs it on a
        side stream underneath the (latency-bound, 1024-token) encoder."""
        cfg, st, D = self.cfg, self.store, self.D
        Nd = B * T
        mel = mel_specs.contiguous()
        melshift = self._empty(Nd, cfg.mel_dim, dtype=BF16)
        ops.shift_cast(mel, melshift.view(B, T, cfg.mel_dim))
        y = self._empty(Nd, D)
        # dropout(dropout(proj, p_in) + PE, p_enc): model.py:525-531 (the PE module is shared with the encoder)
        ctx["drop_in"] = self._ds("dec.in", dcfg.decoder_input if dcfg else 0.0, "dec.pe", p_enc)
        if ctx["drop_in"] is None:
            ops.gemm(melshift, st.w("mel_projection_in.weight"), y, bias=st.p("mel_projection_in.bias"),
                     resid=st.pe[:T], resid_mod=T)
        else:
            t_in = self._empty(Nd, D)
            ops.gemm(melshift, st.w("mel_projection_in.weight"), t_in, bias=st.p("mel_projection_in.bias"))
            # forward needs the two masks separately (the PE is added between them); ctx["drop_in"] (their
            # product) is what the backward applies to the weight-gradient operand
            spec = ops.DropSpec()
            spec.state = self.drop_state.data_ptr()
            spec.site_a, spec.thr_a = self.drop_sites["dec.in"], ops.drop_thr(dcfg.decoder_input)
            spec.site_b, spec.thr_b = self.drop_sites["dec.pe"], ops.drop_thr(p_enc)
            scale_a = 1.0 / ops.drop_keep(dcfg.decoder_input)
            spec.scale = scale_a / ops.drop_keep(p_enc)
            ops.dec_in_drop(t_in, st.pe[:T], y, T, spec, scale_a)
        s1: dict = {}
        pre = "decoder.layers.0."
        y, ln = self._attn_fwd(pre + "self_attn.", y, B, T, pre + "norm1.", True, ctx["mel_pad"], None, T, s1,
                               drop=self._branch_specs("dec.0.self", p_dec, T, False), next_ln=(pre + "norm2.", "bf16"))
        return melshift, (y, ln), s1

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def forward(self, phoneme_indices, mel_specs, phoneme_durations, pitch_targets, energy_targets,
                stress_indices=None, expanded_len: Optional[int] = None,
                text_padding_mask: Optional[torch.Tensor] = None, mel_padding_mask: Optional[torch.Tensor] = None):
        """Training forward.  Returns ((mel, log_dur, stop, pitch, energy), ctx).
        text_padding_mask (B, P) u8, 1 = padding: replaces the default ``phoneme_indices == 0`` (model.py:586-589);
        mel_padding_mask (B, T) u8: key-padding mask of the decoder self-attention (``tgt_key_padding_mask``,
        model.py:648-654; the trainer passes None)."""
        cfg, st, D, H = self.cfg, self.store, self.D, self.H
        B, P = phoneme_indices.shape
        T = mel_specs.shape[1]
        Ne, Nd = B * P, B * T
        va = "duration_adaptor.variance_adaptor."
        ctx: dict = {"B": B, "P": P, "T": T}
        self._live = []          # previous step's tensors may be recycled from here on
        if expanded_len is None:
            # T' = max_b sum(d) (reference utils/lengths.py:44-47 also synchronises here)
            expanded_len = max(1, int(phoneme_durations.clamp(min=0).sum(dim=1).max().item()))
        Tp = int(expanded_len)
        if Tp < 3:
            raise RuntimeError("expanded length < 3 frames is not supported")
        ctx["Tp"] = Tp
        dcfg = self.dropout if self._drop_on else None
        p_enc = dcfg.encoder if dcfg else 0.0
        p_dec = dcfg.decoder if dcfg else 0.0
        if dcfg is not None:      # new RNG step + this step's stochastic-depth factors
            self._path_table = self._empty(len(self._path_rates), B)
            ops.drop_begin(self.drop_state, self._path_site_dev, self._path_p_dev, self._path_table, B)

        ctx["mel_pad"] = mel_padding_mask
        dec_head = None
        if self.multi_stream:                 # decoder input + layer-0 self-attention do not need the encoder
            with self._on("d0"):
                dec_head = self._decoder_head(ctx, mel_specs, B, T, dcfg, p_enc, p_dec)

        # ---- encoder -------------------------------------------------------------------------
        idx = phoneme_indices.contiguous()
        stress = stress_indices.contiguous() if stress_indices is not None else None
        x = self._empty(Ne, D)
        ctx["drop_pe"] = self._ds("enc.pe", p_enc)
        ops.embed_fwd(idx, stress, st.p("text_embedding.weight"), st.p("stress_embedding.weight"), st.pe, x, P,
                      drop=ctx["drop_pe"])
        if text_padding_mask is not None:
            text_pad = text_padding_mask
        else:
            text_pad = self._empty(B, P, dtype=torch.uint8)
            ops.eq_mask(idx, 0, text_pad)
        enc_saved = []
        ln = None                             # LayerNorm of the next sub-layer, produced by the previous sub-layer's tail kernel
        for i in range(cfg.n_encoder_layers):
            pre = f"transformer_encoder_layers.{i}."
            s1, s2 = {}, {}
            x, ln = self._attn_fwd(pre + "self_attn.", x, B, P, pre + "norm1.", False, text_pad, None, P, s1,
                                   drop=self._branch_specs(f"enc.{i}.attn", p_enc, P, False), pre_ln=ln,
                                   next_ln=(pre + "norm2.", "bf16"))
            last = i == cfg.n_encoder_layers - 1
            x, ln = self._ffn_fwd(pre + "ff.", x, pre + "norm2.", cfg.encoder_ff_dim, s2,
                                  drop=self._branch_specs(f"enc.{i}.ffn", p_enc, P, True), pre_ln=ln,
                                  next_ln=("encoder_norm.", "f32") if last else (f"transformer_encoder_layers.{i + 1}.norm1.", "bf16"))
            enc_saved.append((s1, s2))
        enc, enc_mean, enc_rstd = ln          # the final encoder norm (fp32) came out of the last FFN's tail
        ctx.update(idx=idx, stress=stress, text_pad=text_pad, enc_saved=enc_saved, enc_in=x,
                   enc_mean=enc_mean, enc_rstd=enc_rstd)

        # ---- variance adaptor ------------------------------------------------------------------
        gt = self._geom(B, P)
        sv_dur: dict = {}
        with self._on("vp"):
            xg_tok = self._zeros(gt.R + 2, D, dtype=BF16)
            ops.scatter_rows(enc, gt.row_of_tok, xg_tok[1:])
            log_dur = self._vp_fwd(va + "duration_predictor.", xg_tok, gt, text_pad, sv_dur, "duration")

        dur = phoneme_durations.contiguous()
        lr_idx = self._empty(B, Tp, dtype=torch.int32)
        lengths = self._empty(B, dtype=torch.int32)
        ops.lr_index(dur, lr_idx, lengths)
        flags = self._zeros(2, dtype=torch.int32)
        pitch_t, energy_t = pitch_targets.contiguous(), energy_targets.contiguous()
        ops.range_flag(pitch_t, flags[0:1])
        ops.range_flag(energy_t, flags[1:2])
        gf = self._geom(B, Tp)
        xg_frm = self._zeros(gf.R + 2, D, dtype=BF16)
        mem = self._empty(Nd, D, dtype=BF16)
        p_idx = self._empty(B, T, dtype=torch.int32)
        e_idx = self._empty(B, T, dtype=torch.int32)
        fmask_t = self._empty(B, T, dtype=torch.uint8)
        fmask_p = self._empty(B, Tp, dtype=torch.uint8)
        ops.expand_adapt(enc, lr_idx, lengths, pitch_t, energy_t, flags, st.pitch_bins, st.energy_bins,
                         st.p(va + "pitch_embedding.weight"), st.p(va + "energy_embedding.weight"),
                         gf.row_of_tok, xg_frm[1:], mem, p_idx, e_idx, fmask_t, fmask_p, B, P, D, Tp, T)
        if self.spec_spans is not None:      # SpecAugment on the decoder's cross-attention memory only
            ops.spec_augment(mem.view(B, T, D), self.spec_spans, self.spec_n_time, self.spec_n_feat)
        sv_pitch, sv_energy = {}, {}
        with self._on("vp"):
            pitch_pred = self._vp_fwd(va + "pitch_predictor.", xg_frm, gf, fmask_p, sv_pitch, "pitch")
        with self._on("enc"):
            energy_pred = self._vp_fwd(va + "energy_predictor.", xg_frm, gf, fmask_p, sv_energy, "energy")
        ctx.update(sv_dur=sv_dur, sv_pitch=sv_pitch, sv_energy=sv_energy, lr_idx=lr_idx, lengths=lengths,
                   mem=mem, p_idx=p_idx, e_idx=e_idx, fmask_t=fmask_t, fmask_p=fmask_p)

        # ---- decoder ---------------------------------------------------------------------------
        if dec_head is None:
            dec_head = self._decoder_head(ctx, mel_specs, B, T, dcfg, p_enc, p_dec)
        else:
            self._join("d0")
        melshift, (y, ln), s1_first = dec_head
        dec_saved = []
        kv_pre = [None] * cfg.n_decoder_layers
        if self.multi_stream:                 # all six cross-attention K/V projections depend on `mem` only
            with self._on("kv"):
                for i in range(cfg.n_decoder_layers):
                    raw_kv, nkv = self._cross_kv(f"decoder.layers.{i}.cross_attn.", mem, Nd, T)
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    kv_pre[i] = (raw_kv, nkv, ev)
        for i in range(cfg.n_decoder_layers):
            pre = f"decoder.layers.{i}."
            s1, s2, s3 = {}, {}, {}
            if i == 0:
                s1 = s1_first                 # already computed by _decoder_head (with norm2's LayerNorm)
            else:
                y, ln = self._attn_fwd(pre + "self_attn.", y, B, T, pre + "norm1.", True, mel_padding_mask, None, T, s1,
                                       drop=self._branch_specs(f"dec.{i}.self", p_dec, T, False), pre_ln=ln,
                                       next_ln=(pre + "norm2.", "bf16"))
            y, ln = self._attn_fwd(pre + "cross_attn.", y, B, T, pre + "norm2.", False, fmask_t, mem, T, s2,
                                   kv_pre=kv_pre[i], drop=self._branch_specs(f"dec.{i}.cross", p_dec, T, False), pre_ln=ln,
                                   next_ln=(pre + "norm3.", "bf16"))
            last = i == cfg.n_decoder_layers - 1
            y, ln = self._ffn_fwd(pre + "ff.", y, pre + "norm3.", cfg.decoder_ff_dim, s3,
                                  drop=self._branch_specs(f"dec.{i}.ffn", p_dec, T, True), pre_ln=ln,
                                  next_ln=("decoder.norm.", "bf16") if last else (f"decoder.layers.{i + 1}.norm1.", "bf16"))
            dec_saved.append((s1, s2, s3))
        yn, dn_mean, dn_rstd = ln             # decoder.norm came out of the last FFN's tail
        mel_pred = self._empty(Nd, cfg.mel_dim)
        ops.gemm(yn, st.w("mel_projection_out.weight"), mel_pred, bias=st.p("mel_projection_out.bias"))
        stop = self._empty(Nd)
        ops.stop_head_fwd(yn, st.p("stop_token_predictor.weight"), st.p("stop_token_predictor.bias"), stop)
        ctx.update(melshift=melshift, dec_saved=dec_saved, dec_in=y, dn_mean=dn_mean, dn_rstd=dn_rstd, yn=yn)
        self._join_all()
        outs = (mel_pred.view(B, T, cfg.mel_dim), log_dur, stop.view(B, T), pitch_pred, energy_pred)
        return outs, ctx

    # ------------------------------------------------------------------------------------------
    # losses (+ gradients wrt the five outputs)
    # ------------------------------------------------------------------------------------------
    def losses(self, outs, mel_specs, phoneme_durations, stop_targets, pitch_targets, energy_targets,
               mel_lengths, phoneme_lengths, loss_scale: Optional[torch.Tensor] = None):
        mel_pred, log_dur, stop, pitch_pred, energy_pred = outs
        B, T, C = mel_specs.shape
        P = phoneme_durations.shape[1]
        Tp = pitch_pred.shape[1]
        lc = self.loss_cfg
        acc = self._empty(10, dtype=torch.float64)
        losses = self._empty(6)
        g = {"mel": self._empty(B * T, C, dtype=BF16), "dur": self._empty(B, P), "stop": self._empty(B * T),
             "pitch": self._empty(B, Tp), "energy": self._empty(B, Tp)}
        ops.losses_fwd_bwd(mel_pred, mel_specs.contiguous(), log_dur, phoneme_durations.contiguous(), stop,
                           stop_targets.contiguous(), pitch_pred, pitch_targets.contiguous(), energy_pred,
                           energy_targets.contiguous(), mel_lengths.contiguous(), phoneme_lengths.contiguous(),
                           (lc.w_dur, lc.w_stop, lc.w_pitch, lc.w_energy), lc.stop_pos_weight,
                           lc.huber_delta_var, loss_scale, acc, losses, g["mel"], g["dur"], g["stop"],
                           g["pitch"], g["energy"])
        return losses, g

    # ------------------------------------------------------------------------------------------
    # backward: accumulates into store.grads
    # ------------------------------------------------------------------------------------------
    def backward(self, ctx: dict, g: dict):
        for _ in self.backward_parts(ctx, g, None):
            pass

    def early_grad_ranges(self, split_layer: int) -> List[Tuple[int, int]]:
        """Element ranges of the flat gradient buffer that are FINAL when backward_parts(…, split_layer) yields:
        embeddings + encoder + predictors (front of the buffer) and decoder layers >= split_layer + heads (tail).
        The complement — pitch / energy embeddings, mel_projection_in, decoder layers < split_layer — is one range."""
        st = self.store
        a = st.entries["duration_adaptor.variance_adaptor.pitch_embedding.weight"].offset
        b = st.entries[f"decoder.layers.{split_layer}.self_attn.w_q.weight"].offset
        return [(0, a), (b, st.total)]

    def backward_parts(self, ctx: dict, g: dict, split_layer):
        """Backward pass.  split_layer (an int or a list of decoder layer indices, None = off): after the backward of each
        listed decoder layer, `early_reduce_hook(layer)` runs on the "comm" side stream, which first waits for everything
        enqueued so far on every stream — the data-parallel step all-reduces the gradient ranges that are final at that
        point (early_grad_ranges) underneath the rest of the backward.  Without a hook the generator joins all streams
        and yields there instead."""
        cfg, st, D = self.cfg, self.store, self.D
        B, P, T, Tp = ctx["B"], ctx["P"], ctx["T"], ctx["Tp"]
        Ne, Nd = B * P, B * T
        va = "duration_adaptor.variance_adaptor."
        dmel = g["mel"]                                     # [Nd, mel] bf16
        gf, gt = self._geom(B, Tp), self._geom(B, P)
        # (iv) duration loss -> duration predictor -> encoder -> embeddings: disjoint from the decoder
        # backward (the expansion is detached), so it runs on its own stream.
        with self._on("enc"):
            denc = self._vp_bwd(va + "duration_predictor.", g["dur"], gt, ctx["sv_dur"], need_dx=True)
            dx = self._empty(Ne, D)
            dx_bf = self._empty(Ne, D, dtype=BF16)
            ops.layernorm_bwd(denc, ctx["enc_in"], ctx["enc_mean"], ctx["enc_rstd"], st.p("encoder_norm.weight"),
                              None, dx, dx_bf, st.g("encoder_norm.weight"), st.g("encoder_norm.bias"))
            for i in reversed(range(cfg.n_encoder_layers)):
                pre = f"transformer_encoder_layers.{i}."
                s1, s2 = ctx["enc_saved"][i]
                dx, dx_bf = self._ffn_bwd(pre + "ff.", dx, pre + "norm2.", cfg.encoder_ff_dim, s2,
                                          next_drop=s1["drop"]["out"], next_dbias=st.g(pre + "self_attn.w_o.bias"))
                dx, dx_bf = self._attn_bwd(pre + "self_attn.", dx, dx_bf, B, P, pre + "norm1.", False,
                                           ctx["text_pad"], None, P, s1, None, False, bias_done=True)
            ops.embed_bwd(dx, ctx["idx"], ctx["stress"], st.g("text_embedding.weight"),
                          st.g("stress_embedding.weight"), drop=ctx["drop_pe"])
        # (iii) pitch / energy losses -> their predictors only (input is the detached expansion)
        with self._on("vp"):
            self._vp_bwd(va + "pitch_predictor.", g["pitch"], gf, ctx["sv_pitch"], need_dx=False)
            self._vp_bwd(va + "energy_predictor.", g["energy"], gf, ctx["sv_energy"], need_dx=False)
            # (ii) stop loss -> stop head only (it reads a detached decoder output)
            ops.stop_head_bwd(g["stop"], ctx["yn"], st.g("stop_token_predictor.weight"),
                              st.g("stop_token_predictor.bias"))
        # (i) mel loss -> mel head -> decoder -> mel_projection_in and the pitch/energy embedding rows
        self._wgrad(dmel, ctx["yn"], st.g("mel_projection_out.weight"), st.g("mel_projection_out.bias"))
        dyn = self._empty(Nd, D)
        ops.gemm(dmel, st.w("mel_projection_out.weight"), dyn, b_mn_major=True)
        dy = self._empty(Nd, D)
        dy_bf = self._empty(Nd, D, dtype=BF16)
        ops.layernorm_bwd(dyn, ctx["dec_in"], ctx["dn_mean"], ctx["dn_rstd"], st.p("decoder.norm.weight"), None,
                          dy, dy_bf, st.g("decoder.norm.weight"), st.g("decoder.norm.bias"))
        dmem = self._empty(Nd, D)
        first = True
        for i in reversed(range(cfg.n_decoder_layers)):
            pre = f"decoder.layers.{i}."
            s1, s2, s3 = ctx["dec_saved"][i]
            dy, dy_bf = self._ffn_bwd(pre + "ff.", dy, pre + "norm3.", cfg.decoder_ff_dim, s3,
                                      next_drop=s2["drop"]["out"], next_dbias=st.g(pre + "cross_attn.w_o.bias"))
            dy, dy_bf = self._attn_bwd(pre + "cross_attn.", dy, dy_bf, B, T, pre + "norm2.", False, ctx["fmask_t"],
                                       ctx["mem"], T, s2, dmem, first, next_drop=s1["drop"]["out"],
                                       next_dbias=st.g(pre + "self_attn.w_o.bias"), bias_done=True)
            first = False
            # layer 0: the bf16 copy feeds mel_projection_in's weight gradient through both input dropouts
            dy, dy_bf = self._attn_bwd(pre + "self_attn.", dy, dy_bf, B, T, pre + "norm1.", True, ctx["mel_pad"], None, T,
                                       s1, None, False, next_drop=ctx["drop_in"] if i == 0 else None,
                                       next_dbias=st.g("mel_projection_in.bias") if i == 0 else None, bias_done=True)
            if split_layer is not None and i > 0 and (i == split_layer or (isinstance(split_layer, (list, tuple))
                                                                           and i in split_layer)):
                if self.early_reduce_hook is not None and self.multi_stream:
                    # everything in early_grad_ranges(split_layer) has been ENQUEUED by now (encoder / predictor backward
                    # on their streams, weight gradients of layers >= split_layer on w0 / w1 / kv): the comm stream waits
                    # for exactly that work, the main chain goes on with the lower decoder layers without waiting
                    comm = self._side["comm"]
                    cur = torch.cuda.current_stream(self.device)
                    comm.wait_stream(cur)
                    for name in ("enc", "vp", "w0", "w1", "kv"):
                        comm.wait_stream(self._side[name])
                    if comm not in self._forked:
                        self._forked.append(comm)
                    with torch.cuda.stream(comm):
                        self.early_reduce_hook(i)
                else:
                    self._join_all()
                    yield
        self._wgrad(dy_bf, ctx["melshift"], st.g("mel_projection_in.weight"), None)   # bias: layer 0's layernorm_bwd
        self._join_all()
        # memory gradient (accumulated on the "kv" stream) reaches only the pitch / energy embedding rows
        if self.spec_spans is not None:
            ops.spec_augment(dmem.view(B, T, D), self.spec_spans, self.spec_n_time, self.spec_n_feat)
        ops.adapt_bwd(dmem, ctx["p_idx"].view(-1), ctx["e_idx"].view(-1), st.g(va + "pitch_embedding.weight"),
                      st.g(va + "energy_embedding.weight"))

    @staticmethod
    def draw_spec_spans(B: int, T: int, D: int, time_mask_max: int = 5, freq_mask_max: int = 3,
                        num_time_masks: int = 1, num_freq_masks: int = 2) -> torch.Tensor:
        """Host-side span sampling with the SAME torch.randint call sequence as the reference's
        KokoroTrainer._apply_spec_augment (trainer.py:1594-1603): int32 [B, nt + nf, 2] = (start, length)."""
        spans = torch.zeros(B, num_time_masks + num_freq_masks, 2, dtype=torch.int32)
        time_limit = max(1, min(time_mask_max, T // 4))
        for b in range(B):
            for k in range(num_time_masks):
                t = int(torch.randint(0, time_limit, (1,)).item())
                t0 = int(torch.randint(0, max(1, T - t), (1,)).item())
                spans[b, k, 0], spans[b, k, 1] = t0, t
            for k in range(num_freq_masks):
                f = int(torch.randint(0, max(1, freq_mask_max), (1,)).item())
                f0 = int(torch.randint(0, max(1, D - f), (1,)).item())
                spans[b, num_time_masks + k, 0], spans[b, num_time_masks + k, 1] = f0, f
        return spans

    def set_spec_augment(self, spans: Optional[torch.Tensor], n_time: int = 1, n_feat: int = 2) -> None:
        """spans: host or device int32 [B, n_time + n_feat, 2], or None to disable.  The device copy lives in ONE
        buffer per span-table shape that is never re-allocated: captured CUDA graphs (which are keyed by batch shape
        and by whether SpecAugment is on, see TrainStep.stage) keep reading valid memory when batch sizes alternate."""
        if spans is None:
            self.spec_spans = None
            return
        if not hasattr(self, "_spec_bufs"):
            self._spec_bufs: Dict[Tuple[int, ...], torch.Tensor] = {}
        key = tuple(spans.shape)
        buf = self._spec_bufs.get(key)
        if buf is None:
            buf = torch.empty(spans.shape, dtype=torch.int32, device=self.device)
            self._spec_bufs[key] = buf
        if not spans.is_cuda:
            # pinned staging ring: the async copy of step n must not see the host write of step n + 1
            if not hasattr(self, "_spec_ring"):
                self._spec_ring, self._spec_i = {}, 0
            ring = self._spec_ring.get(key)
            if ring is None:
                ring = torch.empty((16,) + key, dtype=torch.int32).pin_memory()
                self._spec_ring[key] = ring
            slot = ring[self._spec_i % 16]
            self._spec_i += 1
            slot.copy_(spans.to(torch.int32))
            spans = slot
        buf.copy_(spans, non_blocking=True)
        self.spec_spans = buf
        self.spec_n_time, self.spec_n_feat = n_time, n_feat

    def zero_grad(self):
        ops.zero_(self.store.grads)
