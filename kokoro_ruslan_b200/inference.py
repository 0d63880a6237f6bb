"""Autoregressive inference on the device — host-side mirror of the reference's
``KokoroModel.forward_inference`` (src/kokoro/model/model.py:675-779), the inference branch of
``VarianceAdaptor.forward`` (model/variance_predictor.py:335-437) and ``KokoroGenerator.generate``
(model/generator.py:24-127).

Two parts:

* ``InferenceEngine.encode_and_expand`` — encoder, duration predictor, ``round(expm1(log_dur))`` durations, length
  regulation, pitch / energy predictors on the expansion, predicted (clamped) values bucketised into the embeddings,
  cross-attention K/V of all decoder layers: the training path's kernels in eval mode, one pass.
* ``DecodeLoop`` — the frame-by-frame decoder.  One step = ``kr_dec_feed`` → per layer [LayerNorm, fused QKV GEMM,
  ``kr_dec_attn`` (cache append + attention), out-projection GEMM with bias + residual, LayerNorm, Q GEMM,
  ``kr_dec_attn`` (cross), out-projection GEMM, LayerNorm, FFN GEMM, GLU, FFN GEMM, RMSNorm + residual] →
  ``kr_dec_finish``.  The frame counter, the cache fill level and the generator's stop rules are device state
  (csrc/kr_decode_core.cuh ``DecState``), every pointer of a step is static, so the step is captured ONCE as a CUDA
  graph and replayed per frame; the host reads the ``done`` flag every ``poll`` frames only (the reference does a
  ``.item()`` per frame, generator.py:67).  The shared kernels run on a 128-row padded batch (rows >= B stay zero), the
  shape they are validated at.

``DecodeLoop`` talks to the kernels through a small backend object; the only backend in the package is
``CudaDecodeBackend`` (libkokoro_b200.so, no fallback).  The CPU test-suite drives the same loop through a backend made of
the kernels' host emulation (tests/test_decode_emu_cpu.py) — orchestration and kernels are checked against the oracle
without a GPU, the -m gpu tests then check the device path.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch

STATE_WORDS = 64
# word offsets inside DecState (csrc/kr_decode_core.cuh)
ST_T, ST_DONE, ST_NFRAMES, ST_LO, ST_HI, ST_EXPECTED, ST_STOP_THR, ST_POST_THR, ST_RING = 0, 1, 2, 3, 4, 5, 6, 7, 8
ROWS = 128            # padded row count of the shared (training-path) kernels during decode


def generation_bounds(expected: int, max_len: int = 4000, min_len_ratio: float = 0.7, min_len_floor: int = 12,
                      max_len_ratio: float = 3.0, max_len_cap: int = 1600) -> Tuple[int, int]:
    """(min_len, max_len) of the generation loop, model.py:737-745."""
    lo = max(min_len_floor, int(expected * min_len_ratio))
    hi = min(max_len, max(expected + 80, int(expected * max_len_ratio)), max_len_cap)
    if hi <= lo:
        hi = min(max_len, lo + 1)
    return lo, hi


def pack_state(lo: int, hi: int, expected: int, stop_thr: float, post_thr: float) -> torch.Tensor:
    """Host image of a fresh DecState (int32 words; the two thresholds are float bit patterns)."""
    st = torch.zeros(STATE_WORDS, dtype=torch.int32)
    st[ST_LO], st[ST_HI], st[ST_EXPECTED] = int(lo), int(hi), int(expected)
    st[ST_STOP_THR:ST_POST_THR + 1] = torch.tensor([stop_thr, post_thr], dtype=torch.float32).view(torch.int32)
    return st


class DecodeLoop:
    """Frame-by-frame decoder over a prepared memory.  ``be`` is the kernel backend (see module docstring)."""

    def __init__(self, be, n_layers: int, D: int, H: int, ff: int, n_mels: int, B: int, Tp: int, t_cap: int,
                 cross_kv: List[torch.Tensor], mem_pad: torch.Tensor):
        if B > 16:
            raise RuntimeError("DecodeLoop decodes at most 16 utterances together")
        assert D == H * 64, "head_dim must be 64"
        self.be, self.L, self.D, self.H, self.ff, self.n_mels = be, n_layers, D, H, ff, n_mels
        self.B, self.Tp, self.t_cap = B, Tp, t_cap
        self.cross_kv, self.mem_pad = cross_kv, mem_pad          # per layer [B * Tp, 2D] bf16 (normalised K | V); [B, Tp] u8
        z = be.zeros
        self.state = z((STATE_WORDS,), torch.int32)
        self.prev = z((B, n_mels), torch.float32)
        self.mel_out = z((B, t_cap, n_mels), torch.float32)
        self.probs = z((t_cap,), torch.float32)
        self.x = [z((ROWS, D), torch.float32) for _ in range(2)]  # residual stream, ping-pong
        self.h = z((ROWS, D), torch.bfloat16)
        self.qkv = z((ROWS, 3 * D), torch.bfloat16)
        self.q = z((ROWS, D), torch.bfloat16)
        self.o = z((ROWS, D), torch.bfloat16)
        self.hff = z((ROWS, 2 * ff), torch.bfloat16)
        self.u = z((ROWS, ff), torch.bfloat16)
        self.yff = z((ROWS, D), torch.float32)
        self.stat = z((2, ROWS), torch.float32)
        self.kc = [z((B, t_cap, D), torch.bfloat16) for _ in range(n_layers)]
        self.vc = [z((B, t_cap, D), torch.bfloat16) for _ in range(n_layers)]
        self.forced: Optional[torch.Tensor] = None
        self._graph = None
        if hasattr(be, "state_for_gemv"):
            be.state_for_gemv = self.state

    # one decode step: every launch reads t / done from self.state, no host-side per-step values
    def step(self) -> None:
        be, D, B = self.be, self.D, self.B
        P, W = be.param, be.weight
        fused = bool(getattr(be, "use_gemv", False)) and B <= 8     # LN / GLU folded into the skinny projections
        cur, nxt = self.x
        be.dec_feed(self.state, self.prev, self.forced, P("mel_projection_in.weight"), P("mel_projection_in.bias"),
                    be.pe, cur, B, D, self.n_mels)
        for i in range(self.L):
            pre = f"decoder.layers.{i}."
            sa, ca, ffp = pre + "self_attn.", pre + "cross_attn.", pre + "ff."
            # self-attention sub-layer (transformers.py:564-569)
            if fused:
                be.ln_gemv(cur, P(pre + "norm1.weight"), P(pre + "norm1.bias"), be.weight_span(sa + "w_q.weight", 3 * D),
                           None, self.qkv, B, glu=False)
            else:
                be.layernorm(cur, P(pre + "norm1.weight"), P(pre + "norm1.bias"), self.h, self.stat)
                be.gemm(self.h, be.weight_span(sa + "w_q.weight", 3 * D), self.qkv, rows=B)
            be.dec_attn(self.state, self.qkv[:, :D], self.qkv[:, D:2 * D], self.qkv[:, 2 * D:], P(sa + "q_norm.weight"),
                        P(sa + "k_norm.weight"), P(sa + "v_norm.weight"), self.kc[i], self.vc[i], -1, None, self.o, B,
                        self.H)
            be.gemm(self.o, W(sa + "w_o.weight"), nxt, bias=P(sa + "w_o.bias"), resid=cur, rows=B)
            cur, nxt = nxt, cur
            # cross-attention sub-layer (:572-578)
            if fused:
                be.ln_gemv(cur, P(pre + "norm2.weight"), P(pre + "norm2.bias"), W(ca + "w_q.weight"), None, self.q, B,
                           glu=False)
            else:
                be.layernorm(cur, P(pre + "norm2.weight"), P(pre + "norm2.bias"), self.h, self.stat)
                be.gemm(self.h, W(ca + "w_q.weight"), self.q, rows=B)
            kv = self.cross_kv[i].view(B, self.Tp, 2 * D)
            be.dec_attn(self.state, self.q, None, None, P(ca + "q_norm.weight"), None, None, kv[:, :, :D], kv[:, :, D:],
                        self.Tp, self.mem_pad, self.o, B, self.H)
            be.gemm(self.o, W(ca + "w_o.weight"), nxt, bias=P(ca + "w_o.bias"), resid=cur, rows=B)
            cur, nxt = nxt, cur
            # GLU feed-forward sub-layer (:581, :105-111)
            if fused:           # LayerNorm + linear1 + GLU in one launch
                be.ln_gemv(cur, P(pre + "norm3.weight"), P(pre + "norm3.bias"), W(ffp + "linear1.weight"),
                           P(ffp + "linear1.bias"), self.u, B, glu=True)
            else:
                be.layernorm(cur, P(pre + "norm3.weight"), P(pre + "norm3.bias"), self.h, self.stat)
                be.gemm(self.h, W(ffp + "linear1.weight"), self.hff, bias=P(ffp + "linear1.bias"), rows=B)
                be.glu(self.hff, self.u)
            be.gemm(self.u, W(ffp + "linear2.weight"), self.yff, bias=P(ffp + "linear2.bias"), rows=B)
            be.rmsnorm_resid(self.yff, P(ffp + "output_norm.weight"), cur, nxt)
            cur, nxt = nxt, cur
        be.dec_finish(self.state, cur, P("decoder.norm.weight"), P("decoder.norm.bias"), P("mel_projection_out.weight"),
                      P("mel_projection_out.bias"), P("stop_token_predictor.weight"), P("stop_token_predictor.bias"),
                      self.mel_out, self.prev, self.probs, B, D, self.n_mels, self.t_cap)

    def run(self, lo: int, hi: int, expected: int, stop_threshold: float = 0.5, post_expected_stop_threshold: float = 0.2,
            forced: Optional[torch.Tensor] = None, poll: int = 32) -> Tuple[torch.Tensor, torch.Tensor]:
        """Generates until the device-side stop rules fire.  Returns (mel (B, n_frames, n_mels) clamped to [-11.5, 2],
        per-frame stop probabilities).  ``forced`` (B, >= hi, n_mels): teacher-forced input frames instead of feedback."""
        if hi > self.t_cap:
            raise ValueError(f"max length {hi} exceeds the cache capacity {self.t_cap}")
        be = self.be
        if forced is not None and (forced.shape[0] != self.B or forced.shape[1] < hi or forced.shape[2] != self.n_mels):
            raise ValueError("forced frames must be (B, >= max length, n_mels)")
        if (forced is None) != (self.forced is None) or (forced is not None and forced.data_ptr() != self.forced.data_ptr()):
            self._graph = None                        # the captured step holds the forced-frame pointer
        self.forced = forced
        be.copy_(self.state, pack_state(lo, hi, expected, stop_threshold, post_expected_stop_threshold))
        be.fill_zero(self.prev)
        if self._graph is None:
            self._graph = be.capture(self.step)
        done = 0
        n_frames = 0
        for _ in range(hi // poll + 2):               # the device stops at t = hi at the latest: a bounded host loop
            for _ in range(poll):
                self._graph()
            st = be.to_host(self.state)
            done, n_frames = int(st[ST_DONE]), int(st[ST_NFRAMES])
            if int(st[ST_T]) > hi:
                raise RuntimeError("decode ran past its bound (corrupt device state)")
            if done:
                break
        if not done:
            raise RuntimeError(f"decode did not finish within {hi} frames (device state: t={int(st[ST_T])})")
        return self.mel_out[:, :n_frames].clone(), self.probs[:n_frames].clone()


class CudaDecodeBackend:
    """libkokoro_b200.so backend of DecodeLoop: validated training-path kernels + the three decode kernels."""

    def __init__(self, engine):
        from . import ops
        from ._lib import check, lib
        self.eng, self.ops, self._check, self._lib = engine, ops, check, lib
        self.st = engine.store
        self.device = engine.device
        self.pe = self.st.pe
        n = int(lib().kr_dec_state_size())
        if n != STATE_WORDS * 4:
            raise RuntimeError(f"DecState is {n} bytes in libkokoro_b200.so, {STATE_WORDS * 4} expected")
        import os
        # projections of the decode step: kr_dec_gemv on the <= 8 live rows (LayerNorm prologue / GLU epilogue fused; every
        # SM streams a slice of the weights once).  Measured on B200 (round 2, tools/decode_bench.py, B=1, 400 frames):
        # 372 us / frame against 625 us with the 128-row padded tcgen05 GEMMs, which remain for 9..16 utterances only
        self.use_gemv = True
        # False = the reference's behaviour (new query rotated as position 0, transformers.py:276-277); True rotates it
        # to its true position like the training forward does (an opt-in fix of that train / inference mismatch)
        self.rotate_query = os.environ.get("KR_DECODE_ROPE_QUERY", "0") == "1"
        self.state_for_gemv = None        # DecodeLoop's state tensor once a loop is bound (lets the kernel skip after `done`)

    def zeros(self, shape, dtype):
        return self.ops.zero_(torch.empty(*shape, dtype=dtype, device=self.device))

    def fill_zero(self, t):
        self.ops.zero_(t)

    def copy_(self, dst, src_host):
        dst.copy_(src_host.pin_memory(), non_blocking=True)

    def to_host(self, t):
        return t.cpu()

    def param(self, name):
        return self.st.p(name)

    def weight(self, name):
        return self.st.w(name)

    def weight_span(self, first, rows):
        return self.st.span(self.st.shadow, first, rows, self.eng.D)

    def layernorm(self, x, g, b, out_bf16, stat):
        self.ops.layernorm_fwd(x, g, b, out_bf16, None, stat[0], stat[1])

    def _gemv(self, x_bf16, x_f32, ln_g, ln_b, w, bias, resid, out, rows, glu):
        o, c = self.ops, ctypes
        x = x_bf16 if x_bf16 is not None else x_f32
        n_out = w.shape[0] // 2 if glu else w.shape[0]
        assert x.stride(1) == 1 and w.is_contiguous() and out.stride(1) == 1 and out.shape[1] == n_out
        self._check(self._lib().kr_dec_gemv(o._ptr(self.state_for_gemv), o._ptr(x_bf16), o._ptr(x_f32),
                                            c.c_longlong(x.stride(0)), o._ptr(ln_g), o._ptr(ln_b), o._ptr(w), o._ptr(bias),
                                            o._ptr(resid), c.c_longlong(0 if resid is None else resid.stride(0)),
                                            o._ptr(out), c.c_longlong(out.stride(0)),
                                            c.c_int(int(out.dtype == torch.float32)), c.c_int(int(glu)), c.c_int(rows),
                                            c.c_int(n_out), c.c_int(w.shape[1]), o._stream()), "kr_dec_gemv")

    def gemm(self, a, w, out, bias=None, resid=None, rows=None):
        """Projection of the step: kr_dec_gemv on the `rows` live rows (at most 8 utterances) — every SM streams a slice of
        the weights once; more rows go through the tcgen05 GEMM on the 128-row padded buffers."""
        if self.use_gemv and rows is not None and rows <= 8 and w.shape[1] % 8 == 0 and w.shape[1] <= 1536:
            self._gemv(a, None, None, None, w, bias, resid, out, rows, False)
            return
        self.ops.gemm(a, w, out, bias=bias, resid=resid)

    def ln_gemv(self, x_f32, ln_g, ln_b, w, bias, out, rows, glu):
        """LayerNorm(x) . W^T (+ bias) in one launch; glu: W = linear1, out = gelu(gate) * lin (<= 8 utterances)."""
        self._gemv(None, x_f32, ln_g, ln_b, w, bias, None, out, rows, glu)

    def glu(self, h, u):
        self.ops.glu_fwd(h, u)

    def rmsnorm_resid(self, y, gain, resid, out):
        self.ops.rmsnorm_resid_fwd(y, gain, resid, out)

    def dec_feed(self, state, prev, forced, w_in, b_in, pe, x, B, D, n_mels):
        o, c = self.ops, ctypes
        self._check(self._lib().kr_dec_feed(o._ptr(state), o._ptr(prev), o._ptr(forced),
                                            c.c_int(0 if forced is None else forced.shape[1]), o._ptr(w_in), o._ptr(b_in),
                                            o._ptr(pe), o._ptr(x), c.c_int(B), c.c_int(D), c.c_int(n_mels), o._stream()),
                    "kr_dec_feed")

    def dec_attn(self, state, q, k_raw, v_raw, gq, gk, gv, kc, vc, n_keys, mask, out, B, H):
        o, c = self.ops, ctypes
        assert q.stride(1) == 1 and out.stride(1) == 1 and kc.stride(2) == 1 and vc.stride(2) == 1
        assert kc.stride(0) == vc.stride(0) and kc.stride(1) == vc.stride(1)
        if mask is not None:
            assert mask.dtype == torch.uint8 and mask.is_contiguous() and mask.shape[1] == n_keys
        ld_kv = 0 if k_raw is None else k_raw.stride(0)
        self._check(self._lib().kr_dec_attn(o._ptr(state), o._ptr(q), c.c_longlong(q.stride(0)), o._ptr(k_raw),
                                            o._ptr(v_raw), c.c_longlong(ld_kv), o._ptr(gq), o._ptr(gk), o._ptr(gv),
                                            o._ptr(self.st.rope_cos), o._ptr(self.st.rope_sin), o._ptr(kc), o._ptr(vc),
                                            c.c_longlong(kc.stride(1)), c.c_longlong(kc.stride(0)), c.c_int(n_keys),
                                            o._ptr(mask), o._ptr(out), c.c_longlong(out.stride(0)), c.c_int(B), c.c_int(H),
                                            c.c_float(0.125), c.c_int(int(self.rotate_query)), o._stream()), "kr_dec_attn")

    def dec_finish(self, state, y, ln_g, ln_b, w_out, b_out, w_stop, b_stop, mel_out, next_frame, probs, B, D, n_mels,
                   t_cap):
        o, c = self.ops, ctypes
        self._check(self._lib().kr_dec_finish(o._ptr(state), o._ptr(y), o._ptr(ln_g), o._ptr(ln_b), o._ptr(w_out),
                                              o._ptr(b_out), o._ptr(w_stop), o._ptr(b_stop), o._ptr(mel_out),
                                              o._ptr(next_frame), o._ptr(probs), c.c_int(B), c.c_int(D), c.c_int(n_mels),
                                              c.c_int(t_cap), o._stream()), "kr_dec_finish")

    def capture(self, fn):
        """The step as a replayable callable: frame 0 runs eagerly (lazy kernel attributes, descriptor caches), the same
        launch sequence is then captured into a CUDA graph that the remaining frames replay."""
        import os
        if os.environ.get("KR_DECODE_GRAPH", "1") == "0":
            return fn
        return _GraphedStep(fn, self.device)


class _GraphedStep:
    """Frame 0 runs eagerly (it warms up lazy kernel attributes and the TMA-descriptor cache, which must not happen
    under capture); the identical launch sequence is then captured once and replayed for every later frame.  Capture
    only records: the device state is not advanced by it."""

    def __init__(self, fn, device):
        self.fn, self.device, self.graph = fn, device, None

    def __call__(self):
        if self.graph is not None:
            self.graph.replay()
            return
        self.fn()
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.fn()
        self.graph = g


class InferenceEngine:
    """``InferenceEngine(acoustic_engine).generate(phoneme_indices, stress_indices)`` — the device path behind
    ``KokoroModel.forward(mel_specs=None)`` / ``forward_inference`` (model.py:675-779)."""

    def __init__(self, engine):
        self.eng = engine
        self.be = CudaDecodeBackend(engine)
        self.last_predictions: dict = {}

    @torch.no_grad()
    def encode_and_expand(self, phoneme_indices: torch.Tensor, stress_indices: Optional[torch.Tensor],
                          durations: Optional[torch.Tensor] = None, text_padding_mask: Optional[torch.Tensor] = None):
        """Inference branch of _encode_and_expand (model.py:450-508).  Returns (mem bf16 [B*Tp, D], frame mask u8 [B, Tp],
        log_dur [B, P], Tp).  ``durations`` (B, P) integer frames per token (new, optional) replaces the predicted
        ``round(expm1(log_dur))`` — duration control, and what the parity tests use to pin the expanded length."""
        from . import ops
        eng = self.eng
        cfg, st, D = eng.cfg, eng.store, eng.D
        was_training = eng.training
        eng.training = False                                    # eval mode: no dropout / stochastic depth
        try:
            B, P = phoneme_indices.shape
            Ne = B * P
            va = "duration_adaptor.variance_adaptor."
            eng._live = []
            idx = phoneme_indices.to(eng.device).contiguous()
            stress = stress_indices.to(eng.device).contiguous() if stress_indices is not None else None
            x = eng._empty(Ne, D)
            ops.embed_fwd(idx, stress, st.p("text_embedding.weight"), st.p("stress_embedding.weight"), st.pe, x, P)
            if text_padding_mask is not None:                   # model.py:701-704: explicit mask instead of indices == 0
                text_pad = text_padding_mask.to(eng.device).to(torch.uint8).contiguous()
            else:
                text_pad = eng._empty(B, P, dtype=torch.uint8)
                ops.eq_mask(idx, 0, text_pad)
            for i in range(cfg.n_encoder_layers):
                pre = f"transformer_encoder_layers.{i}."
                x = eng._attn_fwd(pre + "self_attn.", x, B, P, pre + "norm1.", False, text_pad, None, P, {})
                x = eng._ffn_fwd(pre + "ff.", x, pre + "norm2.", cfg.encoder_ff_dim, {})
            enc = eng._empty(Ne, D)
            ops.layernorm_fwd(x, st.p("encoder_norm.weight"), st.p("encoder_norm.bias"), None, enc, eng._empty(Ne),
                              eng._empty(Ne))
            gt = eng._geom(B, P)
            xg_tok = eng._zeros(gt.R + 2, D, dtype=torch.bfloat16)
            ops.scatter_rows(enc, gt.row_of_tok, xg_tok[1:])
            log_dur = eng._vp_fwd(va + "duration_predictor.", xg_tok, gt, text_pad, {}, "duration")
            if durations is None:
                dur = torch.clamp(torch.round(torch.expm1(log_dur)), min=0).to(torch.int64)  # variance_predictor.py:347
            else:
                dur = durations.to(eng.device, torch.int64).clamp(min=0)
            Tp = max(3, int(dur.sum(dim=1).max().item()))                                    # :357-359 (one host sync)
            lr_idx = eng._empty(B, Tp, dtype=torch.int32)
            lengths = eng._empty(B, dtype=torch.int32)
            ops.lr_index(dur.contiguous(), lr_idx, lengths)
            gf = eng._geom(B, Tp)
            flags = eng._zeros(2, dtype=torch.int32)

            p_bin, e_bin = eng._empty(B, Tp, dtype=torch.int32), eng._empty(B, Tp, dtype=torch.int32)

            def expand(pitch, energy):
                xg = eng._zeros(gf.R + 2, D, dtype=torch.bfloat16)
                mem = eng._empty(B * Tp, D, dtype=torch.bfloat16)
                fm_t, fm_p = eng._empty(B, Tp, dtype=torch.uint8), eng._empty(B, Tp, dtype=torch.uint8)
                ops.expand_adapt(enc, lr_idx, lengths, pitch, energy, flags, st.pitch_bins, st.energy_bins,
                                 st.p(va + "pitch_embedding.weight"), st.p(va + "energy_embedding.weight"), gf.row_of_tok,
                                 xg[1:], mem, p_bin, e_bin, fm_t, fm_p, B, P, D, Tp, Tp)
                return xg, mem, fm_p

            zero = eng._zeros(B, Tp)
            xg_frm, _, fmask = expand(zero, zero)              # pass 1: the predictors' input is the plain expansion
            pitch = eng._vp_fwd(va + "pitch_predictor.", xg_frm, gf, fmask, {}, "pitch")
            energy = eng._vp_fwd(va + "energy_predictor.", xg_frm, gf, fmask, {}, "energy")
            # pass 2: predicted values, clamped to [0, 1] (:410, :431), select the embeddings
            _, mem, fmask = expand(pitch.clamp(0.0, 1.0).contiguous(), energy.clamp(0.0, 1.0).contiguous())
            # the adaptor's inference outputs besides the memory (variance_predictor.py:433-437) stay readable
            self.last_predictions = {"pitch": pitch, "energy": energy, "pitch_bins": p_bin, "energy_bins": e_bin,
                                     "encoder": enc}
            return mem, fmask, log_dur, Tp       # nothing above forks onto the engine's side streams
        finally:
            eng.training = was_training

    @torch.no_grad()
    def generate(self, phoneme_indices: torch.Tensor, stress_indices: Optional[torch.Tensor] = None, max_len: int = 4000,
                 stop_threshold: float = 0.5, post_expected_stop_threshold: float = 0.2, min_len_ratio: float = 0.7,
                 min_len_floor: int = 12, max_len_ratio: float = 3.0, max_len_cap: int = 1600,
                 forced: Optional[torch.Tensor] = None, durations: Optional[torch.Tensor] = None,
                 return_stop_probs: bool = False, text_padding_mask: Optional[torch.Tensor] = None):
        eng = self.eng
        cfg, D = eng.cfg, eng.D
        mem, fmask, _, Tp = self.encode_and_expand(phoneme_indices, stress_indices, durations, text_padding_mask)
        B = phoneme_indices.shape[0]
        lo, hi = generation_bounds(Tp, min(max_len, cfg.max_decoder_seq_len), min_len_ratio, min_len_floor,
                                   max_len_ratio, max_len_cap)
        cross = []
        for i in range(cfg.n_decoder_layers):                   # precompute_cross_attention_kv (model.py:747-752)
            _, nkv = eng._cross_kv(f"decoder.layers.{i}.cross_attn.", mem, B * Tp, Tp)
            cross.append(nkv)
        t_cap = ((hi + 63) // 64) * 64
        loop = DecodeLoop(self.be, cfg.n_decoder_layers, D, eng.H, cfg.decoder_ff_dim, cfg.mel_dim, B, Tp, t_cap, cross,
                          fmask.contiguous())
        mel, probs = loop.run(lo, hi, Tp, stop_threshold, post_expected_stop_threshold, forced=forced)
        return (mel, probs) if return_stop_probs else mel


class Synthesizer:
    """Phoneme indices -> waveform on the device: the acoustic half of ``KokoroTTS.text_to_speech``
    (reference src/kokoro/inference/inference.py:489-560: ``model.forward_inference`` then
    ``VocoderManager.mel_to_audio``) without the text front-end (phonemisation is host string processing and out of this
    path's scope).  ``vocoder`` is a ``kokoro_ruslan_b200.hifigan.HiFiGANGenerator`` (or anything callable on a
    (B, frames, n_mels) CUDA mel)."""

    def __init__(self, model, vocoder, trim_trailing_silence: bool = True):
        self.model, self.vocoder, self.trim = model, vocoder, trim_trailing_silence

    @torch.no_grad()
    def __call__(self, phoneme_indices: torch.Tensor, stress_indices: Optional[torch.Tensor] = None,
                 **generate_kwargs) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (audio (B, samples), mel (B, frames, n_mels)); samples = frames * hop (256).  With
        ``trim_trailing_silence`` (single utterance, like the reference) the mel is cut at the reference's conservative
        trailing-silence point first (inference.py:590-621; one host read of the frame count)."""
        mel = self.model.forward_inference(phoneme_indices, stress_indices=stress_indices, **generate_kwargs)
        if self.trim and mel.shape[0] == 1 and mel.shape[1] > 0:
            from .features import trailing_trim_end
            mel = mel[:, :int(trailing_trim_end(mel)[0])]
        audio = self.vocoder(mel.contiguous())               # (B, T, 80) is auto-detected like the reference's forward
        if audio.dim() == 3:
            audio = audio.squeeze(1)
        return audio, mel
