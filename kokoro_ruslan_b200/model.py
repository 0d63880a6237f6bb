"""``KokoroModel`` — drop-in for the reference module on the training hot path
(src/kokoro/model/model.py: constructor :35-47, ``forward`` :782-818, ``set_memory_augment`` :212-220,
``get_model_info`` :820-845).

Same constructor keywords, same ``forward`` signature and 5-tuple, same ``named_parameters()`` /
``state_dict()`` names and shapes (308 parameters, 311 state-dict keys), so the reference trainer's
name-based optimizer grouping, pre-clipping, EMA and checkpoints work on it unchanged.  The compute
is the CUDA engine: ``forward`` runs the whole training forward as kernel launches and returns
tensors wired into autograd through ONE ``torch.autograd.Function``; ``loss.backward()`` then runs
the hand-scheduled CUDA backward, which accumulates into a flat gradient buffer that every
parameter's ``.grad`` is a view of.  Nothing here computes on the CPU or through ATen math.

Dropout and stochastic depth follow the constructor arguments in ``train()`` mode (fused into the
kernels, counter-based RNG) and are off in ``eval()``.  Reference behaviours that are deliberately
NOT reproduced on this path are listed in DESIGN.md.  Autoregressive inference is
available as ``forward_inference`` (device KV-cache decode, ``inference.py``); ``forward(mel_specs=None)``
keeps raising NotImplementedError until that path has had its first hardware validation run.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

from .engine import AcousticEngine, DropoutConfig
from .params import ModelConfig


class _TrainingForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, model, idx, mel, dur, pitch, energy, stress):
        eng = model.engine
        if model.external_optimizer:
            eng.store.refresh_shadow()           # an external optimizer updates the fp32 masters only
        expanded_len = None
        if not dur.is_cuda:
            expanded_len = max(1, int(dur.clamp(min=0).sum(dim=1).max()))
        dev = eng.device
        args = [t.to(dev, non_blocking=True) if t is not None else None for t in (idx, mel, dur, pitch, energy, stress)]
        if args[3] is None:
            args[3] = torch.zeros(mel.shape[:2], device=dev)
        if args[4] is None:
            args[4] = torch.zeros(mel.shape[:2], device=dev)
        outs, ectx = eng.forward(args[0], args[1].float(), args[2], args[3].float(), args[4].float(), args[5],
                                 expanded_len=expanded_len)
        ctx.model, ctx.ectx = model, ectx
        ctx.shapes = [tuple(o.shape) for o in outs]
        return tuple(outs)

    @staticmethod
    def backward(ctx, dmel, ddur, dstop, dpitch, denergy):
        model, ectx = ctx.model, ctx.ectx
        eng = model.engine
        B, T, C = ctx.shapes[0]

        def z(g, shape):
            return torch.zeros(shape, device=eng.device) if g is None else g.float().contiguous()
        g = {"mel": z(dmel, ctx.shapes[0]).reshape(B * T, C).to(torch.bfloat16).contiguous(),
             "dur": z(ddur, ctx.shapes[1]), "stop": z(dstop, ctx.shapes[2]).reshape(-1),
             "pitch": z(dpitch, ctx.shapes[3]), "energy": z(denergy, ctx.shapes[4])}
        if model._params[0].grad is None:        # first backward after zero_grad(set_to_none=True)
            eng.zero_grad()
        eng.backward(ectx, g)
        for p, gv in zip(model._params, model._grad_views):
            p.grad = gv
        return (None,) * 8


class KokoroModel:
    def __init__(self, vocab_size: int, mel_dim: int = 80, hidden_dim: int = 512, n_encoder_layers: int = 6,
                 n_heads: int = 8, encoder_ff_dim: int = 2048, encoder_dropout: float = 0.1,
                 decoder_dropout: Optional[float] = None, decoder_input_dropout: float = 0.1,
                 n_decoder_layers: int = 6, decoder_ff_dim: int = 2048, max_decoder_seq_len: int = 4000,
                 enable_profiling: bool = False, gradient_checkpointing: bool = True, checkpoint_segments: int = 2,
                 use_variance_predictor: bool = True, variance_filter_size: int = 256, variance_kernel_size: int = 3,
                 variance_dropout: float = 0.1, n_variance_bins: int = 256, pitch_min: float = 50.0,
                 pitch_max: float = 800.0, energy_min: float = 0.0, energy_max: float = 100.0,
                 use_stochastic_depth: bool = True, stochastic_depth_rate: float = 0.1,
                 use_stress_embedding: bool = True, qk_norm: bool = False, ffn_output_norm: bool = True,
                 device=None, seed: int = 0):
        if not use_variance_predictor or not qk_norm or not ffn_output_norm or not use_stress_embedding:
            raise NotImplementedError("the B200 path implements the trainer's configuration: variance predictor, "
                                      "QK-norm, FFN output norm and stress embedding enabled (trainer.py:356-382)")
        if variance_kernel_size != 3:
            raise NotImplementedError("variance predictor kernel size must be 3")
        self.vocab_size, self.mel_dim, self.hidden_dim = vocab_size, mel_dim, hidden_dim
        self.max_decoder_seq_len = max_decoder_seq_len
        self.use_variance_predictor = True
        self.gradient_checkpointing, self.checkpoint_segments = gradient_checkpointing, checkpoint_segments
        self.enable_profiling = enable_profiling
        self.dropouts = dict(encoder=encoder_dropout, decoder=decoder_dropout, decoder_input=decoder_input_dropout,
                             variance=variance_dropout, stochastic_depth=stochastic_depth_rate if use_stochastic_depth else 0.0)
        cfg = ModelConfig(vocab_size=vocab_size, mel_dim=mel_dim, hidden_dim=hidden_dim,
                          n_encoder_layers=n_encoder_layers, n_heads=n_heads, encoder_ff_dim=encoder_ff_dim,
                          n_decoder_layers=n_decoder_layers, decoder_ff_dim=decoder_ff_dim,
                          max_decoder_seq_len=max_decoder_seq_len, variance_filter_size=variance_filter_size,
                          n_variance_bins=n_variance_bins)
        dec_p = decoder_dropout if decoder_dropout is not None else encoder_dropout     # model.py:78
        self.engine = AcousticEngine(cfg, device, with_ema=False, dropout=DropoutConfig(
            encoder=encoder_dropout, decoder=dec_p, decoder_input=decoder_input_dropout, variance=variance_dropout,
            stochastic_depth=stochastic_depth_rate if use_stochastic_depth else 0.0, seed=seed))
        self.engine.store.init_default(seed=seed)
        self.training = True
        self.external_optimizer = True
        self._memory_augment_fn: Optional[Callable] = None
        st = self.engine.store
        self._names: List[str] = list(st.order)
        self._params = [torch.nn.Parameter(st.ref_view(st.params, n), requires_grad=True) for n in self._names]
        self._grad_views = [st.ref_view(st.grads, n) for n in self._names]
        self._hook = torch.zeros(1, device=self.engine.device, requires_grad=True)
        # attribute shims the reference trainer / checkpoint manager read
        self.transformer_encoder_layers = [None] * n_encoder_layers
        self.decoder = type("DecoderShim", (), {"num_layers": n_decoder_layers, "layers": [None] * n_decoder_layers})()

    # ---- nn.Module-like surface ------------------------------------------------------------------
    def named_parameters(self) -> Iterator[Tuple[str, torch.nn.Parameter]]:
        return iter(zip(self._names, self._params))

    def parameters(self) -> Iterator[torch.nn.Parameter]:
        return iter(self._params)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self.engine.store.ordered_state_dict()

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        self.engine.store.load_state_dict(sd, strict=strict)
        return self

    def train(self, mode: bool = True):
        self.training = bool(mode)
        self.engine.training = self.training      # eval(): dropout / stochastic depth off (nn.Module semantics)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("KokoroModel (B200) is CUDA-only: there is no CPU / MPS fallback")
        return self

    def zero_grad(self, set_to_none: bool = True):
        for p in self._params:
            p.grad = None
        if not set_to_none:
            self.engine.zero_grad()
            for p, gv in zip(self._params, self._grad_views):
                p.grad = gv

    @property
    def variance_adaptor(self):
        return self

    def get_model_info(self) -> dict:
        total = sum(p.numel() for p in self._params)
        return {"vocab_size": self.vocab_size, "mel_dim": self.mel_dim, "hidden_dim": self.hidden_dim,
                "n_encoder_layers": len(self.transformer_encoder_layers), "n_decoder_layers": self.decoder.num_layers,
                "total_parameters": total, "trainable_parameters": total, "model_size_mb": total * 4 / (1024 * 1024),
                "gradient_checkpointing": {"enabled": False, "segments": self.checkpoint_segments,
                                           "memory_savings_estimated": "n/a (activations are kept: 180 GB HBM)"}}

    # ---- augmentation hook -------------------------------------------------------------------------
    def set_memory_augment(self, fn: Optional[Callable]) -> None:
        """Reference hook (model.py:212-220).  ``None`` disables.  A callable is accepted for API
        compatibility but the masking itself runs in ``kr_spec_augment``: pass the span table through
        ``set_spec_augment_spans`` (``AcousticEngine.draw_spec_spans`` draws it with the reference's RNG
        call sequence).  An arbitrary Python callable cannot be traced into the CUDA path."""
        self._memory_augment_fn = fn
        if fn is None:
            self.engine.set_spec_augment(None)

    def set_spec_augment_spans(self, spans: Optional[torch.Tensor], n_time: int = 1, n_feat: int = 2) -> None:
        self.engine.set_spec_augment(spans, n_time, n_feat)

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, phoneme_indices: torch.Tensor, mel_specs: Optional[torch.Tensor] = None,
                phoneme_durations: Optional[torch.Tensor] = None, stop_token_targets: Optional[torch.Tensor] = None,
                pitch_targets: Optional[torch.Tensor] = None, energy_targets: Optional[torch.Tensor] = None,
                text_padding_mask: Optional[torch.Tensor] = None, mel_padding_mask: Optional[torch.Tensor] = None,
                stress_indices: Optional[torch.Tensor] = None):
        if mel_specs is None:
            import os
            if os.environ.get("KR_FORWARD_INFERENCE", "0") == "1":      # the reference's dispatch, model.py:813-818
                return self.forward_inference(phoneme_indices, max_len=self.max_decoder_seq_len,
                                              text_padding_mask=text_padding_mask, stress_indices=stress_indices)
            raise NotImplementedError("forward(mel_specs=None): the device decode path (forward_inference, SURVEY.md 8(f) "
                                      "N2) has not had its first hardware validation run; call forward_inference() "
                                      "directly or set KR_FORWARD_INFERENCE=1")
        if phoneme_durations is None or stop_token_targets is None:
            raise ValueError("phoneme_durations and stop_token_targets must be provided for training mode.")
        if text_padding_mask is not None or mel_padding_mask is not None:
            raise NotImplementedError("explicit padding masks: the trainer always passes None (text mask = "
                                      "indices == 0, no mel mask; trainer.py:3226-3230)")
        outs = _TrainingForward.apply(self._hook, self, phoneme_indices, mel_specs, phoneme_durations, pitch_targets,
                                      energy_targets, stress_indices)
        return outs

    __call__ = forward

    def forward_inference(self, phoneme_indices: torch.Tensor, max_len: int = 4000, stop_threshold: float = 0.5,
                          text_padding_mask: Optional[torch.Tensor] = None, min_len_ratio: float = 0.7,
                          min_len_floor: int = 12, max_len_ratio: float = 3.0, max_len_cap: int = 1600,
                          post_expected_stop_threshold: float = 0.2,
                          stress_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reference signature (model/model.py:675-687).  Returns the generated mel (B, n_frames, mel_dim) clamped to
        [-11.5, 2].  Like the reference it does not change train / eval mode; the decode itself always runs without
        dropout (the reference's generator is only ever used under ``eval()``)."""
        if text_padding_mask is not None:
            raise NotImplementedError("explicit text padding mask (the reference derives it as indices == 0, model.py:704)")
        from .inference import InferenceEngine
        if getattr(self, "_inference", None) is None:
            self._inference = InferenceEngine(self.engine)
        return self._inference.generate(phoneme_indices, stress_indices, max_len=max_len, stop_threshold=stop_threshold,
                                        post_expected_stop_threshold=post_expected_stop_threshold,
                                        min_len_ratio=min_len_ratio, min_len_floor=min_len_floor,
                                        max_len_ratio=max_len_ratio, max_len_cap=max_len_cap)
