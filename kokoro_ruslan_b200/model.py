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
dispatches to it like the reference's ``forward`` (model.py:813-818).

The class IS a ``torch.nn.Module``: its sub-module tree mirrors the reference's module names (name-space nodes that own
views into the engine's flat buffers), so ``named_modules()`` / ``modules()`` / ``copy.deepcopy`` / attribute access —
everything the reference ``KokoroTrainer`` does to its model (trainer.py:835, 845-881, 2049-2055) — work on it.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterator, List, Optional, Tuple

import torch

from .engine import AcousticEngine, DropoutConfig
from .params import ModelConfig


class _TrainingForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, model, idx, mel, dur, pitch, energy, stress, text_pad=None, mel_pad=None):
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
        masks = [m.to(dev).to(torch.uint8).contiguous() if m is not None else None for m in (text_pad, mel_pad)]
        outs, ectx = eng.forward(args[0], args[1].float(), args[2], args[3].float(), args[4].float(), args[5],
                                 expanded_len=expanded_len, text_padding_mask=masks[0], mel_padding_mask=masks[1])
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
        return (None,) * 10


class _Node(torch.nn.Module):
    """Inner node of the module tree that mirrors the reference's module NAMES (``decoder.layers.3.ff.linear1`` ...).
    It owns nothing but the Parameters / buffers registered on it — views into the engine's flat buffers — so that
    ``named_modules()``, ``named_parameters()``, ``modules()`` and attribute access (``m.decoder.layers[3].ff.linear1.weight``,
    trainer.py:845-881) behave as on the reference model.  It is never called."""

    def forward(self, *a, **k):          # pragma: no cover - compute lives in the CUDA engine
        raise RuntimeError("sub-modules of the B200 KokoroModel are name-space nodes: call the model itself")

    def __getitem__(self, i):            # ModuleList-style access for the numeric children (layers[3])
        return self._modules[str(i)]

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())


class KokoroModel(torch.nn.Module):
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
        super().__init__()
        if not use_variance_predictor or not qk_norm or not ffn_output_norm or not use_stress_embedding:
            raise NotImplementedError("the B200 path implements the trainer's configuration: variance predictor, "
                                      "QK-norm, FFN output norm and stress embedding enabled (trainer.py:356-382)")
        if variance_kernel_size != 3:
            raise NotImplementedError("variance predictor kernel size must be 3")
        # constructor arguments, kept for copy.deepcopy (the reference trainer's EMA model, trainer.py:835)
        self._ctor = dict(vocab_size=vocab_size, mel_dim=mel_dim, hidden_dim=hidden_dim, n_encoder_layers=n_encoder_layers,
                          n_heads=n_heads, encoder_ff_dim=encoder_ff_dim, encoder_dropout=encoder_dropout,
                          decoder_dropout=decoder_dropout, decoder_input_dropout=decoder_input_dropout,
                          n_decoder_layers=n_decoder_layers, decoder_ff_dim=decoder_ff_dim,
                          max_decoder_seq_len=max_decoder_seq_len, enable_profiling=enable_profiling,
                          gradient_checkpointing=gradient_checkpointing, checkpoint_segments=checkpoint_segments,
                          use_variance_predictor=use_variance_predictor, variance_filter_size=variance_filter_size,
                          variance_kernel_size=variance_kernel_size, variance_dropout=variance_dropout,
                          n_variance_bins=n_variance_bins, pitch_min=pitch_min, pitch_max=pitch_max,
                          energy_min=energy_min, energy_max=energy_max, use_stochastic_depth=use_stochastic_depth,
                          stochastic_depth_rate=stochastic_depth_rate, use_stress_embedding=use_stress_embedding,
                          qk_norm=qk_norm, ffn_output_norm=ffn_output_norm, seed=seed)
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
        # plain attribute (not a sub-module): nn.Module.__setattr__ only intercepts Modules / Parameters / buffers
        self.engine = AcousticEngine(cfg, device, with_ema=False, dropout=DropoutConfig(
            encoder=encoder_dropout, decoder=dec_p, decoder_input=decoder_input_dropout, variance=variance_dropout,
            stochastic_depth=stochastic_depth_rate if use_stochastic_depth else 0.0, seed=seed))
        self.engine.store.init_default(seed=seed)
        self.external_optimizer = True
        self._memory_augment_fn: Optional[Callable] = None
        st = self.engine.store
        self._names: List[str] = list(st.order)
        self._params = [torch.nn.Parameter(st.ref_view(st.params, n), requires_grad=True) for n in self._names]
        self._grad_views = [st.ref_view(st.grads, n) for n in self._names]
        self._hook = torch.zeros(1, device=self.engine.device, requires_grad=True)
        self._build_module_tree()

    def _build_module_tree(self) -> None:
        """Registers every Parameter under the reference's dotted name (and the three buffers where the reference has
        them), creating ``_Node`` sub-modules on first use.  The flat names are in the reference's registration (DFS)
        order, so ``named_parameters()`` reproduces it — asserted below, the optimizer grouping depends on it."""
        st = self.engine.store

        def node_for(path: List[str]) -> torch.nn.Module:
            m: torch.nn.Module = self
            for part in path:
                if part not in m._modules:
                    m.add_module(part, _Node())
                m = m._modules[part]
            return m
        for name, p in zip(self._names, self._params):
            *path, leaf = name.split(".")
            node_for(path).register_parameter(leaf, p)
        node_for(["positional_encoding"]).register_buffer("pe", st.pe.unsqueeze(0))
        va = node_for(["duration_adaptor", "variance_adaptor"])
        va.register_buffer("pitch_bins", st.pitch_bins)
        va.register_buffer("energy_bins", st.energy_bins)
        self.decoder.num_layers = len(self.decoder.layers)            # read by the reference's checkpoint manager
        assert [n for n, _ in torch.nn.Module.named_parameters(self)] == self._names

    # ---- nn.Module surface: storage is the engine's flat buffers ------------------------------------
    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        """311 keys in the reference's order; every value is a VIEW of the flat fp32 buffer (the reference trainer's EMA
        update mutates ``ema_model.state_dict()`` tensors in place, trainer.py:1506-1515 — that works on these views)."""
        sd = self.engine.store.ordered_state_dict()
        out = destination if destination is not None else type(sd)()
        for k, v in sd.items():
            out[prefix + k] = v
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True, assign: bool = False):
        self.engine.store.load_state_dict(sd, strict=strict)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    def train(self, mode: bool = True):
        super().train(mode)
        self.engine.training = bool(mode)         # eval(): dropout / stochastic depth off (nn.Module semantics)
        return self

    def to(self, *args, **kwargs):
        device = kwargs.get("device", args[0] if args else None)
        if isinstance(device, (str, torch.device)) and torch.device(device).type != "cuda":
            raise RuntimeError("KokoroModel (B200) is CUDA-only: there is no CPU / MPS fallback")
        return self

    def _apply(self, fn, recurse: bool = True):   # .cuda() / .float() / .half(): the flat buffers never move or change type
        return self

    def zero_grad(self, set_to_none: bool = True):
        for p in self._params:
            p.grad = None
        if not set_to_none:
            self.engine.zero_grad()
            for p, gv in zip(self._params, self._grad_views):
                p.grad = gv

    def __deepcopy__(self, memo):
        """``copy.deepcopy(model)`` (the reference trainer builds its EMA model this way, trainer.py:835): a new model
        with its own engine / flat buffers on the same device, same weights, same train / eval mode."""
        twin = type(self)(device=self.engine.device, **self._ctor)
        twin.engine.store.load_state_dict(self.engine.store.state_dict())
        twin.train(self.training)
        twin.external_optimizer = self.external_optimizer
        memo[id(self)] = twin
        return twin

    @property
    def variance_adaptor(self):
        """model.py:222-231: the wrapped VarianceAdaptor (here: its name-space node with the parameters and bins)."""
        return self.duration_adaptor.variance_adaptor

    def get_model_info(self) -> dict:
        """model.py:820-845."""
        total = sum(p.numel() for p in self._params)
        return {"vocab_size": self.vocab_size, "mel_dim": self.mel_dim, "hidden_dim": self.hidden_dim,
                "n_encoder_layers": len(self.transformer_encoder_layers), "n_decoder_layers": self.decoder.num_layers,
                "total_parameters": total, "trainable_parameters": total, "model_size_mb": total * 4 / (1024 * 1024),
                "gradient_checkpointing": {"enabled": False, "segments": self.checkpoint_segments,
                                           "memory_savings_estimated": "n/a (activations are kept: 180 GB HBM)"}}

    # ---- augmentation hook -------------------------------------------------------------------------
    def set_memory_augment(self, fn: Optional[Callable]) -> None:
        """Reference hook (model.py:212-220): ``fn`` maps the (B, T, D) cross-attention memory to its augmented version,
        ``None`` disables.  The masking itself runs in ``kr_spec_augment``, so a callable is TRANSLATED, every training
        forward, into the span table that kernel reads: ``fn`` is applied to a (B, T, D) probe of ones on the host and the
        zero pattern it leaves must be, per utterance, a union of whole time rows and whole feature columns — which is
        what the reference trainer's closure over ``_apply_spec_augment`` produces (trainer.py:1578-1604, 2049-2055), with
        its own ``torch.randint`` draws.  Any other callable (values other than 0 / 1, patterns that are not row / column
        bands) raises instead of being silently ignored."""
        self._memory_augment_fn = fn
        if fn is None:
            self.engine.set_spec_augment(None)

    def _translate_memory_augment(self, B: int, T: int) -> None:
        """Probe the user's callable and load the equivalent span table (see set_memory_augment)."""
        D = self.hidden_dim
        probe = torch.ones(B, T, D, dtype=torch.float32)
        out = self._memory_augment_fn(probe)
        if not torch.is_tensor(out) or tuple(out.shape) != (B, T, D):
            raise NotImplementedError("set_memory_augment: the callable must return a (B, T, D) tensor")
        out = out.detach().cpu()
        zero = out == 0
        if not bool(((out == 1) | zero).all()):
            raise NotImplementedError("set_memory_augment: only zero-masking augmentations can be translated to "
                                      "kr_spec_augment spans (the callable changed values to something other than 0)")
        rows, cols = zero.all(dim=2), zero.all(dim=1)                      # (B, T), (B, D)
        if not torch.equal(zero, rows.unsqueeze(2) | cols.unsqueeze(1)):
            raise NotImplementedError("set_memory_augment: the mask is not a union of whole time rows and whole "
                                      "feature columns per utterance (SpecAugment bands)")

        def runs(mask_1d):
            m = mask_1d.to(torch.int8)
            d = torch.diff(torch.cat([torch.zeros(1, dtype=torch.int8), m, torch.zeros(1, dtype=torch.int8)]))
            starts, ends = (d == 1).nonzero().flatten(), (d == -1).nonzero().flatten()
            return [(int(a), int(b - a)) for a, b in zip(starts, ends)]
        tr, fr = [runs(rows[b]) for b in range(B)], [runs(cols[b]) for b in range(B)]
        nt, nf = max(1, max(len(r) for r in tr)), max(1, max(len(r) for r in fr))
        spans = torch.zeros(B, nt + nf, 2, dtype=torch.int32)
        for b in range(B):
            for k, (a, n) in enumerate(tr[b]):
                spans[b, k, 0], spans[b, k, 1] = a, n
            for k, (a, n) in enumerate(fr[b]):
                spans[b, nt + k, 0], spans[b, nt + k, 1] = a, n
        self.engine.set_spec_augment(spans, nt, nf)

    def set_spec_augment_spans(self, spans: Optional[torch.Tensor], n_time: int = 1, n_feat: int = 2) -> None:
        self.engine.set_spec_augment(spans, n_time, n_feat)

    # ---- forward -------------------------------------------------------------------------------------
    def forward(self, phoneme_indices: torch.Tensor, mel_specs: Optional[torch.Tensor] = None,
                phoneme_durations: Optional[torch.Tensor] = None, stop_token_targets: Optional[torch.Tensor] = None,
                pitch_targets: Optional[torch.Tensor] = None, energy_targets: Optional[torch.Tensor] = None,
                text_padding_mask: Optional[torch.Tensor] = None, mel_padding_mask: Optional[torch.Tensor] = None,
                stress_indices: Optional[torch.Tensor] = None):
        if mel_specs is None:                                            # the reference's dispatch, model.py:813-818
            return self.forward_inference(phoneme_indices, max_len=self.max_decoder_seq_len,
                                          text_padding_mask=text_padding_mask, stress_indices=stress_indices)
        if phoneme_durations is None or stop_token_targets is None:
            raise ValueError("phoneme_durations and stop_token_targets must be provided for training mode.")
        if self.training and self._memory_augment_fn is not None:        # model.py:636-639
            dsum = int(phoneme_durations.clamp(min=0).sum(dim=1).max()) if phoneme_durations.numel() else 0
            if max(1, dsum) >= 3:
                self._translate_memory_augment(mel_specs.shape[0], mel_specs.shape[1])
        outs = _TrainingForward.apply(self._hook, self, phoneme_indices, mel_specs, phoneme_durations, pitch_targets,
                                      energy_targets, stress_indices, text_padding_mask, mel_padding_mask)
        return outs

    def forward_inference(self, phoneme_indices: torch.Tensor, max_len: int = 4000, stop_threshold: float = 0.5,
                          text_padding_mask: Optional[torch.Tensor] = None, min_len_ratio: float = 0.7,
                          min_len_floor: int = 12, max_len_ratio: float = 3.0, max_len_cap: int = 1600,
                          post_expected_stop_threshold: float = 0.2,
                          stress_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reference signature (model/model.py:675-687).  Returns the generated mel (B, n_frames, mel_dim) clamped to
        [-11.5, 2].  Like the reference it does not change train / eval mode; the decode itself always runs without
        dropout (the reference's generator is only ever used under ``eval()``)."""
        from .inference import InferenceEngine
        if getattr(self, "_inference", None) is None:
            self._inference = InferenceEngine(self.engine)
        return self._inference.generate(phoneme_indices, stress_indices, max_len=max_len, stop_threshold=stop_threshold,
                                        text_padding_mask=text_padding_mask,
                                        post_expected_stop_threshold=post_expected_stop_threshold,
                                        min_len_ratio=min_len_ratio, min_len_floor=min_len_floor,
                                        max_len_ratio=max_len_ratio, max_len_cap=max_len_cap)
