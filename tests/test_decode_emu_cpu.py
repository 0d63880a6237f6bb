"""Autoregressive decode on the CPU through the PRODUCT's DecodeLoop (kokoro_ruslan_b200/inference.py: buffer wiring,
device-state protocol, polling) with a test backend made of
  * the decode kernels' own bodies compiled as a host emulation (csrc/kr_decode_core.cuh, tests/emu/decode_emu.cpp), and
  * plain torch restatements of the shared training-path kernels (LayerNorm, bf16 GEMM with fp32 accumulation, GLU,
    RMSNorm + residual) in the same mixed precision,
against the oracle's forward_inference, which is pinned to the live reference (tests/test_inference_cpu.py).
The device path (CudaDecodeBackend) is checked by tests/test_inference_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BF16 = torch.bfloat16


@pytest.fixture(scope="module", params=["seq", "simt"])
def emu(request, tmp_path_factory):
    """The kernels' bodies as a host library: "seq" = one sequential thread per block; "simt" = one host thread per CUDA
    thread with the real block / warp geometry, barriers and warp reductions (tests/emu/emu_simt.h)."""
    simt = request.param == "simt"
    so = tmp_path_factory.mktemp("emu") / ("decode_emu_%s.so" % request.param)
    flags = ["-DKR_HOST_EMU_SIMT", "-pthread", "-I", os.path.join(HERE, "emu")] if simt else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *flags, "-x", "c++", "-I",
                    os.path.join(ROOT, "kokoro_ruslan_b200", "csrc"), os.path.join(HERE, "emu", "decode_emu.cpp"), "-o", str(so)],
                   check=True)
    lib = ctypes.CDLL(str(so))
    lib.simt = simt
    return lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


class EmuBackend:
    """Test double of CudaDecodeBackend: same interface, CPU tensors."""

    def __init__(self, lib, sd, max_len):
        from kokoro_ruslan_b200.params import rope_tables
        self.lib, self.sd = lib, sd
        self.pe = sd["positional_encoding.pe"][0].contiguous()
        self.cos, self.sin = rope_tables(max_len, 64)
        self._w = {}
        self.launches = 0
        self.rotate_query = False

    def zeros(self, shape, dtype):
        return torch.zeros(*shape, dtype=dtype)

    def fill_zero(self, t):
        t.zero_()

    def copy_(self, dst, src):
        dst.copy_(src)

    def to_host(self, t):
        return t.clone()

    def param(self, name):
        return self.sd[name].contiguous()

    def weight(self, name):
        if name not in self._w:
            self._w[name] = self.sd[name].to(BF16).contiguous()
        return self._w[name]

    def weight_span(self, first, rows):
        pre = first[:-len("w_q.weight")]
        key = first + "|span"
        if key not in self._w:
            self._w[key] = torch.cat([self.sd[pre + f"w_{c}.weight"] for c in "qkv"]).to(BF16).contiguous()
        assert self._w[key].shape[0] == rows
        return self._w[key]

    def layernorm(self, x, g, b, out_bf16, stat):
        out_bf16.copy_(F.layer_norm(x, (x.shape[-1],), g, b, 1e-5).to(BF16))

    def gemm(self, a, w, out, bias=None, resid=None, rows=None):
        y = a.float() @ w.float().t()
        if bias is not None:
            y = y + bias
        if resid is not None:
            y = y + resid
        out.copy_(y.to(out.dtype))

    def glu(self, h, u):
        gate, lin = h.float().chunk(2, dim=-1)
        u.copy_((F.gelu(gate) * lin).to(BF16))

    def rmsnorm_resid(self, y, gain, resid, out):
        eps = torch.finfo(torch.float32).eps
        out.copy_(resid + y * torch.rsqrt(y.pow(2).mean(dim=-1, keepdim=True) + eps) * gain)

    def dec_feed(self, state, prev, forced, w_in, b_in, pe, x, B, D, n_mels):
        self.launches += 1
        assert self.lib.emu_dec_feed(_p(state), _p(prev), _p(forced), 0 if forced is None else forced.shape[1], _p(w_in),
                                     _p(b_in), _p(pe), _p(x), B, D, n_mels) == 0

    def dec_attn(self, state, q, k_raw, v_raw, gq, gk, gv, kc, vc, n_keys, mask, out, B, H):
        self.launches += 1
        ll = ctypes.c_longlong
        assert kc.stride(2) == 1 and kc.stride() == vc.stride()
        assert self.lib.emu_dec_attn(_p(state), _p(q), ll(q.stride(0)), _p(k_raw), _p(v_raw),
                                     ll(0 if k_raw is None else k_raw.stride(0)), _p(gq), _p(gk), _p(gv), _p(self.cos),
                                     _p(self.sin), _p(kc), _p(vc), ll(kc.stride(1)), ll(kc.stride(0)), n_keys, _p(mask),
                                     _p(out), ll(out.stride(0)), B, H, ctypes.c_float(0.125), int(self.rotate_query)) == 0

    def dec_finish(self, state, y, ln_g, ln_b, w_out, b_out, w_stop, b_stop, mel_out, next_frame, probs, B, D, n_mels,
                   t_cap):
        self.launches += 1
        assert self.lib.emu_dec_finish(_p(state), _p(y), _p(ln_g), _p(ln_b), _p(w_out), _p(b_out), _p(w_stop), _p(b_stop),
                                       _p(mel_out), _p(next_frame), _p(probs), B, D, n_mels, t_cap) == 0

    def capture(self, fn):
        return fn


class EmuGemvBackend(EmuBackend):
    """Projections through the emulated kr_dec_gemv (the KR_DECODE_GEMV=1 variant of CudaDecodeBackend), with its
    LayerNorm prologue and GLU epilogue."""
    use_gemv = True

    def _gemv(self, x_bf16, x_f32, ln_g, ln_b, w, bias, resid, out, rows, glu):
        self.launches += 1
        ll = ctypes.c_longlong
        x = x_bf16 if x_bf16 is not None else x_f32
        n_out = w.shape[0] // 2 if glu else w.shape[0]
        assert rows <= 8 and w.is_contiguous() and out.shape[1] == n_out
        assert self.lib.emu_dec_gemv(None, _p(x_bf16), _p(x_f32), ll(x.stride(0)), _p(ln_g), _p(ln_b), _p(w), _p(bias),
                                     _p(resid), ll(0 if resid is None else resid.stride(0)), _p(out), ll(out.stride(0)),
                                     int(out.dtype == torch.float32), int(glu), rows, n_out, w.shape[1]) == 0

    def gemm(self, a, w, out, bias=None, resid=None, rows=None):
        self._gemv(a, None, None, None, w, bias, resid, out, rows, False)

    def ln_gemv(self, x_f32, ln_g, ln_b, w, bias, out, rows, glu):
        self._gemv(None, x_f32, ln_g, ln_b, w, bias, None, out, rows, glu)

    def layernorm(self, *a):
        raise AssertionError("the fused variant must not launch a separate LayerNorm")

    def glu(self, *a):
        raise AssertionError("the fused variant must not launch a separate GLU")


def _setup():
    from oracle import acoustic as oa
    f = np.load(os.path.join(HERE, "golden", "inference.npz"))
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=2, n_decoder_layers=2, ff_dim=256,
                            variance_filter=64, max_len=1200)
    sd = oa.seeded_state_dict(cfg, seed=int(f["seed"]))
    sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([float(f["dur_bias"])])
    sd["stop_token_predictor.bias"] = torch.tensor([float(f["stop_bias"])])
    return f, cfg, sd


def _memory(sd, cfg, idx, stress):
    """Oracle encode_and_expand + the normalised cross-attention K | V per layer in the engine's [B*Tp, 2D] layout."""
    from oracle import acoustic as oa
    from oracle import inference as oi
    with torch.no_grad():
        mem, mem_pad, _ = oi.encode_and_expand(sd, cfg, idx, stress)
        B, Tp, D = mem.shape
        H = cfg.n_heads
        mem_bf = mem.to(BF16).float()                                   # the engine's memory is bf16
        cross = []
        for i in range(cfg.n_decoder_layers):
            pre = f"decoder.layers.{i}.cross_attn."
            k = oa._rms((mem_bf @ sd[pre + "w_k.weight"].to(BF16).float().t()).view(B, Tp, H, 64), sd[pre + "k_norm.weight"])
            v = oa._rms((mem_bf @ sd[pre + "w_v.weight"].to(BF16).float().t()).view(B, Tp, H, 64), sd[pre + "v_norm.weight"])
            cross.append(torch.cat([k.reshape(B * Tp, D), v.reshape(B * Tp, D)], dim=1).to(BF16).contiguous())
    return cross, mem_pad.to(torch.uint8).contiguous(), Tp


def _loop(emu, sd, cfg, idx, stress, backend=None):
    from kokoro_ruslan_b200.inference import DecodeLoop, generation_bounds
    cross, mem_pad, Tp = _memory(sd, cfg, idx, stress)
    lo, hi = generation_bounds(Tp)
    t_cap = ((hi + 63) // 64) * 64
    be = (backend or EmuBackend)(emu, sd, cfg.max_len)
    loop = DecodeLoop(be, cfg.n_decoder_layers, cfg.hidden_dim, cfg.n_heads, cfg.ff_dim, cfg.mel_dim, idx.shape[0], Tp,
                      t_cap, cross, mem_pad)
    return loop, be, lo, hi, Tp


def test_state_layout_matches_the_kernel_struct(emu):
    from kokoro_ruslan_b200 import inference as inf
    assert emu.emu_dec_state_size() == inf.STATE_WORDS * 4
    st = inf.pack_state(12, 300, 100, 0.5, 0.2)
    assert st[inf.ST_LO] == 12 and st[inf.ST_HI] == 300 and st[inf.ST_EXPECTED] == 100
    assert st[inf.ST_STOP_THR:inf.ST_POST_THR + 1].view(torch.float32).tolist() == pytest.approx([0.5, 0.2])


def test_teacher_forced_decode_matches_oracle_frame_by_frame(emu):
    """Feed the ORACLE's own output frames: every step then sees exact inputs, so the comparison isolates one step's
    numerics (bf16 operands / caches) from autoregressive drift.  Also pins cache append + RoPE positions."""
    from oracle import inference as oi
    f, cfg, sd = _setup()
    idx, stress = torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"])
    want, want_p, raw = oi.forward_inference(sd, cfg, idx, stress, return_raw=True)
    loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, stress)
    n = want.shape[1]
    forced = torch.zeros(1, hi, cfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]                                     # input of frame t = UN-clamped output t-1
    got, probs = loop.run(lo, hi, Tp, forced=forced, poll=7)
    assert got.shape[1] == n, (got.shape, want.shape)                   # same stop decision on the same inputs
    err = float((got - want).abs().max()) / float(want.abs().max())
    assert err < 1e-2, err
    assert float((probs - torch.tensor(want_p)).abs().max()) < 2e-2


def test_free_running_generation_matches_oracle(emu):
    from oracle import inference as oi
    f, cfg, sd = _setup()
    idx, stress = torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"])
    want = oi.forward_inference(sd, cfg, idx, stress)
    loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, stress)
    got, probs = loop.run(lo, hi, Tp)
    assert abs(got.shape[1] - want.shape[1]) <= 2, (got.shape, want.shape)
    n = min(got.shape[1], want.shape[1])
    err = float((got[:, :n] - want[:, :n]).abs().max()) / float(want.abs().max())
    assert err < 3e-2, err
    assert be.launches == (2 + 2 * cfg.n_decoder_layers) * (-(-got.shape[1] // 32) * 32)   # polling granularity 32


def test_padded_batch_and_stop_threshold(emu):
    """Batch of two with a padded tail (cross-attention key mask, batch-mean stop probability).  Teacher-forced for the
    numerics gate; free-running, this tiny random-weight decoder amplifies the bf16 rounding from frame to frame
    (0.7 % at frame 0 -> 10 % at frame 13), so only the first frames and the frame count are gated there."""
    from oracle import inference as oi
    f, cfg, sd = _setup()
    idx = torch.from_numpy(f["idx2"])
    want, want_p, raw = oi.forward_inference(sd, cfg, idx, None, stop_threshold=0.45, return_raw=True)
    loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, None)
    n = want.shape[1]
    forced = torch.zeros(2, hi, cfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]
    got, probs = loop.run(lo, hi, Tp, stop_threshold=0.45, forced=forced)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) / float(want.abs().max()) < 1e-2
    assert float((probs - torch.tensor(want_p)).abs().max()) < 1e-2
    got, _ = loop.run(lo, hi, Tp, stop_threshold=0.45)                   # same loop object, free-running this time
    assert abs(got.shape[1] - n) <= 2, (got.shape, want.shape)
    assert float((got[:, :3] - want[:, :3]).abs().max()) / float(want.abs().max()) < 3e-2


def test_stop_rules_on_the_device_state(emu):
    """kr_dec_finish's rules in isolation: hard stop at hi, no stop before lo, silence rule after 30 quiet frames."""
    from kokoro_ruslan_b200 import inference as inf
    D, M, B, cap = 128, 80, 2, 64
    z = torch.zeros
    g, b0 = torch.ones(D), z(D)
    w_out, w_stop = z(M, D), z(1, D)
    y = torch.randn(B, D)

    def run(lo, hi, b_out_val, b_stop_val, steps):
        st = inf.pack_state(lo, hi, 40, 0.5, 0.2)
        mel, nxt, probs = z(B, cap, M), z(B, M), z(cap)
        for _ in range(steps):
            assert emu.emu_dec_finish(_p(st), _p(y), _p(g), _p(b0), _p(w_out), _p(torch.full((M,), b_out_val)), _p(w_stop),
                                      _p(torch.tensor([b_stop_val])), _p(mel), _p(nxt), _p(probs), B, D, M, cap) == 0
        return st, mel, nxt, probs

    st, mel, nxt, probs = run(12, 20, -3.0, -5.0, 30)            # never confident: runs to hi, extra steps are no-ops
    assert int(st[inf.ST_DONE]) == 1 and int(st[inf.ST_NFRAMES]) == 20 and int(st[inf.ST_T]) == 20
    assert float(mel[:, :20].max()) == -3.0 and float(mel[:, 20:].abs().max()) == 0.0
    st, *_ = run(12, 60, -3.0, 5.0, 30)                          # confident from the start: stops at the first t >= lo
    assert int(st[inf.ST_NFRAMES]) == 13
    st, mel, nxt, _ = run(12, 60, -12.0, -5.0, 40)               # silence: mean of 30 frames < -9.5 once 30 exist
    assert int(st[inf.ST_NFRAMES]) == 30
    assert float(mel[:, :30].min()) == -11.5 and float(nxt.min()) == -12.0      # output clamped, feedback not


def test_gemv_projection_variant_matches_the_gemm_path(emu):
    """kr_dec_gemv (emulated) in place of the projections: same frames as the torch bf16-GEMM stand-in of the default path
    (both accumulate in fp32 over bf16 operands; only the summation order differs), teacher-forced, batch of two."""
    from oracle import inference as oi
    f, cfg, sd = _setup()
    idx = torch.from_numpy(f["idx2"])
    want, want_p, raw = oi.forward_inference(sd, cfg, idx, None, stop_threshold=0.45, return_raw=True)
    n = want.shape[1]
    outs = []
    for backend in (EmuBackend, EmuGemvBackend):
        loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, None, backend)
        forced = torch.zeros(2, hi, cfg.mel_dim)
        forced[:, 1:n] = raw[:, :n - 1]
        got, _ = loop.run(lo, hi, Tp, stop_threshold=0.45, forced=forced)
        assert got.shape == want.shape
        outs.append(got)
    assert float((outs[1] - want).abs().max()) / float(want.abs().max()) < 1e-2
    # vs the un-fused stand-in: the GLU input is no longer rounded to bf16 in between, so the two differ by about one
    # bf16 rounding of the FFN hidden state (0.7e-2 of max here); both sit within 1e-2 of the fp32 oracle
    assert float((outs[1] - outs[0]).abs().max()) / float(want.abs().max()) < 1e-2
    print("fused / un-fused error vs oracle:", float((outs[1] - want).abs().max()), float((outs[0] - want).abs().max()))
    assert be.launches == (2 + 8 * cfg.n_decoder_layers) * (-(-n // 32) * 32)      # 8 decode-kernel launches per layer


def test_query_rotation_fix_makes_decode_equal_the_causal_training_forward(emu):
    """KR_DECODE_ROPE_QUERY=1 (rotate the new query to its true position): the teacher-forced KV-cache decode then equals
    the TRAINING-style causal pass of the decoder over the same memory — the reference's decode (position 0) does not.
    Isolates the quirk: same kernels, same cache, one flag."""
    import torch.nn.functional as F
    from oracle import acoustic as oa
    from oracle import inference as oi
    f, cfg, sd = _setup()
    idx, stress = torch.from_numpy(f["idx"]), torch.from_numpy(f["stress"])
    with torch.no_grad():
        mem, mem_pad, _ = oi.encode_and_expand(sd, cfg, idx, stress)
        n = 24
        g = torch.Generator().manual_seed(1)
        frames = torch.randn(1, n, cfg.mel_dim, generator=g) * 1.5 - 4.0                 # the "previous frames" fed in
        shifted = F.pad(frames[:, :-1], (0, 0, 1, 0))                                      # teacher forcing: frame t sees t-1
        y = shifted @ sd["mel_projection_in.weight"].t() + sd["mel_projection_in.bias"] + sd["positional_encoding.pe"][0, :n]
        for i in range(cfg.n_decoder_layers):
            y = oa.decoder_block(sd, f"decoder.layers.{i}.", cfg, y, mem, mem_pad)
        y = oa._ln(sd, "decoder.norm.", y)
        causal = (y @ sd["mel_projection_out.weight"].t() + sd["mel_projection_out.bias"]).clamp(-11.5, 2.0)
    outs = {}
    for rotate in (False, True):
        loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, stress)
        be.rotate_query = rotate
        forced = torch.zeros(1, hi, cfg.mel_dim)
        forced[:, :n] = shifted
        got, _ = loop.run(n, n + 1, Tp, forced=forced)                                     # exactly n + 1 frames
        outs[rotate] = got[:, :n]
    scale = float(causal.abs().max())
    assert float((outs[True] - causal).abs().max()) / scale < 1e-2                         # fixed decode == training pass
    assert float((outs[False][:, :1] - causal[:, :1]).abs().max()) / scale < 1e-2          # frame 0: position 0 either way
    assert float((outs[False][:, 4:] - causal[:, 4:]).abs().max()) / scale > 3e-2          # the reference's mismatch


def test_all_padding_memory_gives_a_zero_cross_attention_context(emu):
    """A single phoneme whose predicted duration rounds to zero: the 3-frame minimum memory is ALL padding
    (variance_predictor.py:357-359), torch's SDPA gives the query zero weights and the reference decodes on
    (tests/test_inference_cpu.py pins the oracle to the installed reference on exactly this case).  The decode kernel's
    own source (phase C of dec_attn_body: L > 0 ? o / L : 0) must do the same — no NaN, same frames, same stop step."""
    from oracle import acoustic as oa
    from oracle import inference as oi
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=2, ff_dim=128, variance_filter=64,
                            max_len=1200)
    sd = oa.seeded_state_dict(cfg, seed=6)
    sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([1.2])
    sd["stop_token_predictor.bias"] = torch.tensor([-1.0])
    g = torch.Generator().manual_seed(31)
    torch.randint(1, cfg.vocab_size, (3, 10), generator=g)             # same draw order as the reference-side test
    idx = torch.randint(1, cfg.vocab_size, (1, 1), generator=g)
    mem, mem_pad, _ = oi.encode_and_expand(sd, cfg, idx, None)
    assert mem.shape[1] == 3 and bool(mem_pad.all())                    # the case this test is about
    want, want_p, raw = oi.forward_inference(sd, cfg, idx, None, return_raw=True)
    assert torch.isfinite(want).all() and want.shape[1] == 13           # what the installed reference generates
    loop, be, lo, hi, Tp = _loop(emu, sd, cfg, idx, None)
    n = want.shape[1]
    forced = torch.zeros(1, hi, cfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]
    got, probs = loop.run(lo, hi, Tp, forced=forced, poll=5)
    assert torch.isfinite(got).all() and torch.isfinite(probs).all()
    assert got.shape[1] == n, (got.shape, want.shape)
    assert float((got - want).abs().max()) / float(want.abs().max()) < 1e-2
    assert float((probs - torch.tensor(want_p)).abs().max()) < 2e-2


@pytest.mark.parametrize("label,dur_bias,stop_bias,shape,kw", [
    ("stop never fires", 0.8, -30.0, (1, 6), {}),
    ("stop fires at once", 0.8, 30.0, (1, 6), {}),
    ("small max_len", 1.5, -30.0, (1, 12), dict(max_len=20)),
], ids=["never", "at once", "small max_len"])
def test_generation_bounds_on_the_kernel_source(emu, label, dur_bias, stop_bias, shape, kw):
    """The stop rules of dec_finish_body at the edges the reference-side test pins the oracle on: a stop head that never
    fires ends at the upper bound, one that fires at once ends right after the lower bound, a small max_len caps the run.
    Teacher-forced with the oracle's frames; the kernel source must stop on the same step."""
    from oracle import acoustic as oa
    from oracle import inference as oi
    from kokoro_ruslan_b200.inference import DecodeLoop, generation_bounds
    cfg = oa.AcousticConfig(hidden_dim=128, n_heads=2, n_encoder_layers=1, n_decoder_layers=2, ff_dim=128, variance_filter=64,
                            max_len=1200)
    sd = oa.seeded_state_dict(cfg, seed=6)
    sd["duration_adaptor.variance_adaptor.duration_predictor.linear.bias"] = torch.tensor([dur_bias])
    sd["stop_token_predictor.bias"] = torch.tensor([stop_bias])
    idx = torch.randint(1, cfg.vocab_size, shape, generator=torch.Generator().manual_seed(17))
    want, want_p, raw = oi.forward_inference(sd, cfg, idx, None, return_raw=True, **kw)
    cross, mem_pad, Tp = _memory(sd, cfg, idx, None)
    lo, hi = generation_bounds(Tp, **kw)
    assert (lo, hi) == oi.generation_bounds(Tp, **kw)
    n = want.shape[1]
    assert n == (hi if stop_bias < 0 else lo + 1), (label, n, lo, hi)   # the edge this case is about
    t_cap = ((hi + 63) // 64) * 64
    loop = DecodeLoop(EmuBackend(emu, sd, cfg.max_len), cfg.n_decoder_layers, cfg.hidden_dim, cfg.n_heads, cfg.ff_dim,
                      cfg.mel_dim, idx.shape[0], Tp, t_cap, cross, mem_pad)
    forced = torch.zeros(shape[0], hi, cfg.mel_dim)
    forced[:, 1:n] = raw[:, :n - 1]
    got, probs = loop.run(lo, hi, Tp, forced=forced, poll=9)
    assert got.shape[1] == n, (label, got.shape, want.shape)
    assert float((got - want).abs().max()) / float(want.abs().max()) < 1e-2
