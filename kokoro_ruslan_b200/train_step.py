"""One optimizer step of the acoustic model on one GPU (or one rank of a data-parallel job).

Restates the reference's per-step order of operations (SURVEY.md §8 "Step algorithm";
reference src/kokoro/training/trainer.py:2218-2294 adaptive stabilisation, :3181-3315
``_execute_training_step``, training/runtime_policies.py:14-87 optimizer step, trainer.py:1519-1575
scheduler) around the CUDA engine:

    host batch (pinned) --H2D--> static device buffers
      -> [CUDA graph A]  zero_grad, forward, fused losses (+ grads of the outputs), backward
      -> [collective]    kr_allreduce_sqnorm (world_size > 1 only): all-reduce(SUM) of the flat gradient buffer over
                         NVSwitch symmetric memory, fused with the per-chunk gradient norms and the global clip
      -> [CUDA graph B]  grad norms -> step control -> fused clip+AdamW+EMA -> weight-norm projection
      -> losses[6] stay on the device; the caller decides when to read them.

There are no host synchronisations inside a step: the expanded length T' = max_b sum(d) and the
long-sequence stabiliser are computed from the HOST copy of the batch before the H2D copy, the
finite guards live in the device-side step control (non-finite gradients skip the step).
Graphs are cached per (B, P, T, T', SpecAugment on/off) key; unseen shapes run eagerly once (warm-up) and are then
captured.
"""
from __future__ import annotations

import math
import os
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from ._lib import launch_count
from .engine import AcousticEngine, DropoutConfig, LossConfig
from .optim import FusedAdamW, OptimConfig
from .params import ModelConfig

BATCH_KEYS = ("phoneme_indices", "stress_indices", "phoneme_durations", "mel_specs", "pitches", "energies",
              "stop_token_targets", "mel_lengths", "phoneme_lengths")
PACKED_KEYS = tuple(k for k in BATCH_KEYS if k != "mel_specs")      # the small fields: one packed H2D copy per step
PACK_RING = 4                                                       # pinned host images in flight (guarded by events)


@dataclass
class ScheduleConfig:
    """reference training/config.py:34-35,79-81 + trainer.py:691-775"""
    total_steps: int = 100000          # optimizer steps in the run (epochs * steps/epoch)
    use_warmup: bool = True
    warmup_steps: int = 1200
    warmup_start_lr_ratio: float = 0.01
    max_lr_multiplier: float = 1.0
    pct_start: float = 0.20
    final_div_factor: float = 1.0e4


class WarmupOneCycle:
    """Per-group learning rate for optimizer step k, reproducing the reference's hand-rolled linear
    warm-up followed by torch OneCycleLR(cos, two-phase) (trainer.py:691-775, 1519-1575).

    Reference quirk kept: the LR the optimizer holds *before* the first scheduler call is
    OneCycleLR's initial LR (max_lr / div_factor), so optimizer step 0 runs at the full base LR and
    the warm-up ramp only starts with step 1.
    """

    def __init__(self, base_lr: float, mults: List[float], cfg: ScheduleConfig):
        self.cfg = cfg
        self.base_lr = base_lr
        self.mults = list(mults)
        total = max(1, int(cfg.total_steps))
        warm = int(cfg.warmup_steps) if cfg.use_warmup else 0
        if warm >= total:                       # _apply_warmup_guard, trainer.py:1638-1651
            warm = max(0, total - 1)
        self.warmup_steps = warm
        self.onecycle_steps = max(1, total - warm)
        self.max_lr = base_lr * cfg.max_lr_multiplier
        self.div_factor = max(1.0, float(cfg.max_lr_multiplier)) if cfg.use_warmup else 25.0
        self.warmup_start = base_lr * cfg.warmup_start_lr_ratio
        self.warmup_target = min(base_lr, self.max_lr)
        self.current_optimizer_step = 0         # scheduler calls made so far
        self.sched_step = 0                     # OneCycleLR.last_epoch

    def _onecycle(self, step: int) -> float:
        """Base (mult = 1) OneCycleLR value at `step` (torch two-phase cosine)."""
        total, pct = self.onecycle_steps, self.cfg.pct_start
        initial = self.max_lr / self.div_factor
        min_lr = initial / self.cfg.final_div_factor
        end1 = float(pct * total) - 1.0
        end2 = float(total) - 1.0

        def cos_anneal(start, end, p):
            return end + (start - end) / 2.0 * (math.cos(math.pi * p) + 1.0)

        if step <= end1 or end1 >= end2:
            p = step / end1 if end1 > 0 else 1.0
            return cos_anneal(initial, self.max_lr, min(1.0, p))
        p = (step - end1) / (end2 - end1)
        return cos_anneal(self.max_lr, min_lr, min(1.0, p))

    def lrs(self) -> List[float]:
        """LR per group for the NEXT optimizer step."""
        k = self.current_optimizer_step
        if k == 0:
            base = self._onecycle(0)
        elif self.cfg.use_warmup and (k - 1) < self.warmup_steps:
            prog = (k - 1) / self.warmup_steps
            base = self.warmup_start + (self.warmup_target - self.warmup_start) * prog
        else:
            base = self._onecycle(self.sched_step)
        return [base * m for m in self.mults]

    def advance(self) -> None:
        """Called after a successful optimizer step (runtime_policies.py:81-85)."""
        k = self.current_optimizer_step
        if self.cfg.use_warmup and k < self.warmup_steps:
            pass
        elif self.sched_step < self.onecycle_steps - 1:
            self.sched_step += 1
        self.current_optimizer_step += 1

    def rewind(self, n: int) -> None:
        """Un-does the last n advance() calls (optimizer steps the device-side guard skipped: the reference does not step
        its scheduler on a skipped step, runtime_policies.py:81-85)."""
        k = max(0, self.current_optimizer_step - max(0, int(n)))
        self.current_optimizer_step, self.sched_step = 0, 0
        for _ in range(k):
            self.advance()

    def state_dict(self):
        return {"current_optimizer_step": self.current_optimizer_step, "sched_step": self.sched_step}

    def load_state_dict(self, sd):
        self.current_optimizer_step = int(sd["current_optimizer_step"])
        self.sched_step = int(sd["sched_step"])


def adaptive_stabilisation(T: int, max_dur: int, base_clip: float, frame_thr: float = 1400.0,
                           dur_thr: float = 150.0) -> Tuple[float, float]:
    """(loss scale, clip norm) for long utterances, reference trainer.py:2218-2255: with
    r = max(T/1400, max(d)/150) > 1 the loss is scaled by max(0.25, 1/r) and the clip BECOMES
    max(0.05, 0.5/sqrt(r)) — assigned, not min-ed with the configured clip (:2246-2247; the soft stage of :2233-2236 has
    the same thresholds and is always overwritten by it)."""
    r = max(T / frame_thr, max_dur / dur_thr)
    if r <= 1.0:
        return 1.0, base_clip
    return max(0.25, 1.0 / r), max(0.05, 0.5 / math.sqrt(r))


def validate_batch(batch) -> None:
    """Error conventions of the reference's `_transfer_batch_to_device` (trainer.py:1262-1297): a missing key is a
    KeyError, a None field a ValueError, a non-tensor field a TypeError.  This path needs all nine tensors (the
    variance predictors and the stress embedding are always on, trainer.py:356-382), so the three fields the
    reference treats as optional are required here as well."""
    missing = [k for k in BATCH_KEYS if k not in batch]
    if missing:
        raise KeyError(f"Batch is missing required keys: {missing}")
    for k in BATCH_KEYS:
        v = batch[k]
        if v is None:
            raise ValueError(f"Batch field '{k}' is None")
        if not torch.is_tensor(v):
            raise TypeError(f"Batch field '{k}' must be a tensor, got {type(v).__name__}")
    B = batch["mel_specs"].shape[0]
    for k in BATCH_KEYS:
        if batch[k].shape[0] != B:
            raise ValueError(f"Batch field '{k}' has batch size {batch[k].shape[0]}, expected {B}")


@dataclass
class _Staged:
    """Static device buffers + graph for one batch shape."""
    dev: Dict[str, torch.Tensor]
    # one graph per variant: index 1 = zero the gradient buffer first (start of an accumulation window),
    # index 0 = accumulate onto it
    # (index bit 1 = the closing micro-batch of a data-parallel window: the early all-reduce is part of the graph)
    graph: List[Optional[torch.cuda.CUDAGraph]] = field(default_factory=lambda: [None] * 4)
    graph_losses: List[Optional[torch.Tensor]] = field(default_factory=lambda: [None] * 4)
    losses: Optional[torch.Tensor] = None
    launches: int = 0
    warm: List[int] = field(default_factory=lambda: [0] * 4)
    last_used: int = 0
    # every batch field except the mel lives in ONE device allocation (dev[k] are views at 256-byte offsets), filled from a
    # ring of pinned host images by ONE copy: a step's host batch costs two H2D transfers (this + the mel) instead of nine
    pack_dev: Optional[torch.Tensor] = None
    pack_host: Optional[torch.Tensor] = None                       # [PACK_RING, bytes] pinned
    pack_views: List[Dict[str, torch.Tensor]] = field(default_factory=list)
    pack_done: List[Optional[torch.cuda.Event]] = field(default_factory=list)
    pack_i: int = 0


class TrainStep:
    def __init__(self, model_cfg: Optional[ModelConfig] = None, optim_cfg: Optional[OptimConfig] = None,
                 sched_cfg: Optional[ScheduleConfig] = None, loss_cfg: Optional[LossConfig] = None,
                 device=None, use_graphs: bool = True, process_group=None, max_seq_cap: int = 2000,
                 max_cached_shapes: int = 16, dropout: Optional[DropoutConfig] = None):
        self.engine = AcousticEngine(model_cfg or ModelConfig(), device, with_ema=True, loss_cfg=loss_cfg,
                                     dropout=dropout)
        self.device = self.engine.device
        self.opt = FusedAdamW(self.engine.store, optim_cfg or OptimConfig())
        self.sched = WarmupOneCycle(self.opt.cfg.learning_rate, self.opt.lr_mult, sched_cfg or ScheduleConfig())
        self.use_graphs = use_graphs
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
        # data parallel: the ONE collective of a step is a hand-written kernel over NVSwitch peer / multicast memory that
        # also produces the optimizer's gradient norms and the global clip (parallel.SymmetricGradReducer); there is no
        # second communication path — a box without symmetric-memory support fails here, loudly
        self.split_layer: Optional[int] = None
        self.reducer = None
        self.comm = "none"
        if self.world > 1:
            from .parallel import SymmetricGradReducer, symmetric_memory_usable
            if not symmetric_memory_usable(process_group, self.device):
                raise RuntimeError("kokoro_ruslan_b200: data-parallel training needs CUDA symmetric memory (peer access "
                                   "between all ranks of the node); it is not available on this box")
            self.reducer = SymmetricGradReducer(self.engine.store, self.opt, process_group)
            self.comm = "fused"
            # overlap: the gradient ranges that are final once the backward has passed decoder layer `split_layer`
            # (embeddings + encoder + predictors at the front of the flat buffer, the upper decoder layers + heads at its
            # tail: ~72 % of the bytes) are reduced from the engine's "comm" side stream underneath the lower decoder
            # layers, by a small grid that shares the SMs with the backward kernels; the rest follows graph A
            n_dec = self.engine.cfg.n_decoder_layers
            # measured (round 2, bench shape): 8 GPUs / multicast 6.79 -> 6.55 ms per step with a 32-block early launch
            # (16 and 64 blocks: 6.60); 2 GPUs / peer loads 6.74 -> 7.03 ms (the latency-bound peer path needs ~4 blocks per
            # SM and then takes the SMs away from the backward), so the overlap is used on the multicast path only
            if n_dec >= 2 and self.engine.multi_stream and self.reducer.multicast:
                # two early launches: after decoder layer n/2 (embeddings + encoder + predictors + upper decoder layers +
                # heads, ~72 % of the bytes) and after layer 1 (the layers in between); layer 0, mel_projection_in and the
                # pitch / energy embedding rows (~10 %) follow graph A together with the global clip
                splits = sorted({n_dec // 2, 1}, reverse=True) if n_dec >= 4 else [n_dec // 2]
                starts = self.opt.chunk_start.cpu().tolist()
                ent = self.engine.store.entries
                ca = starts.index(ent["duration_adaptor.variance_adaptor.pitch_embedding.weight"].offset)
                hi = self.opt.n_chunks
                self._early_ranges = {}
                for k, layer in enumerate(splits):
                    cb = starts.index(ent[f"decoder.layers.{layer}.self_attn.w_q.weight"].offset)
                    self._early_ranges[layer] = (0, ca, cb, hi) if k == 0 else (cb, hi, 0, 0)
                    hi = cb
                self._late_ranges = (ca, hi, 0, 0)
                self.split_layer = splits
                self._early_on = False
                self.engine.early_reduce_hook = self._early_reduce
        self.max_seq_cap = max_seq_cap
        # every cached batch shape pins its static input buffers and (once captured) a graph with ~4 GB of
        # activations at the bench shape: dynamic batching produces many shapes, so the cache is LRU-bounded
        self.max_cached_shapes = max_cached_shapes
        self._tick = 0
        # Dynamic batching (DynamicFrameBatchSampler re-packs its buckets every epoch) produces a new (B, P, T, T') shape almost
        # every step: a graph captured for a shape that never comes back costs a device synchronisation, ~0.3 s of capture
        # and, at eviction, the release of its ~4 GB pool — measured 24 - 52 ms per step against 14.5 ms for plain eager
        # launches (tools/dynamic_bench.py).  So capturing is tied to the shape cache's hit rate: always during the first
        # CAPTURE_PROBE steps (fixed-shape runs capture at once), afterwards only while at least CAPTURE_MIN_HIT of the
        # steps arrive with a shape that is already cached.
        self._shape_hits = 0
        self._staged: Dict[Tuple[int, int, int, int, bool], _Staged] = {}
        self._opt_graph: Optional[torch.cuda.CUDAGraph] = None
        self._opt_warm = 0
        self._opt_launches = 0
        # [loss scale / world, clip norm] for the current step: one pinned ring slot -> one H2D copy
        # (+ the sequence id of the step, which travels back with the losses: train_step_host())
        self._scal_ring = torch.zeros(64, 3, dtype=torch.float32).pin_memory()
        self._scal_i = 0
        self._scal_dev = torch.tensor([1.0, self.opt.cfg.max_grad_norm, 0.0], dtype=torch.float32, device=self.device)
        self.loss_scale = self._scal_dev[0:1]
        self.clip = self._scal_dev[1:2]
        # losses[6] + the step's sequence id, written by a D2H copy that is part of the step right after the loss kernel (side
        # stream "l"): the host can read a step's losses while its backward pass and optimizer are still running
        self._loss_host = torch.zeros(8, dtype=torch.float32).pin_memory()
        self._seq = 0
        # The mel batch is read by the first forward kernel and by the loss kernel only.  Once train_step_host() has SEEN a
        # step's losses on the host, that step is past its last reader, so the next batch's mel (2 MB, 40 us of PCIe — the one
        # large input) is copied from a side stream while the backward pass and optimizer of the step still run.
        self._mel_idle = False
        self._copy_stream = (torch.cuda.Stream(device=self.device)
                             if self.device.type == "cuda" and self.engine.multi_stream else None)   # KR_STREAMS=0: one stream
        self._mel_copied: Optional[torch.cuda.Event] = None
        self.launches_last_step = 0
        self.h2d_bytes_last_step = 0

    # ------------------------------------------------------------------------------------------
    @property
    def store(self):
        return self.engine.store

    def load_state_dict(self, sd, strict: bool = True):
        self.engine.store.load_state_dict(sd, strict=strict)

    def state_dict(self):
        return self.engine.store.ordered_state_dict()

    # ------------------------------------------------------------------------------------------
    def _cap(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """_cap_batch_sequence_dimensions, trainer.py:3365-3411 (cap T and P to 2000)."""
        cap = self.max_seq_cap
        T = batch["mel_specs"].shape[1]
        P = batch["phoneme_indices"].shape[1]
        if T <= cap and P <= cap:
            return batch
        out = dict(batch)
        if T > cap:
            for k in ("mel_specs", "pitches", "energies", "stop_token_targets"):
                out[k] = batch[k][:, :cap].contiguous()
            out["mel_lengths"] = batch["mel_lengths"].clamp(max=cap)
        if P > cap:
            for k in ("phoneme_indices", "stress_indices", "phoneme_durations"):
                out[k] = batch[k][:, :cap].contiguous()
            out["phoneme_lengths"] = batch["phoneme_lengths"].clamp(max=cap)
        return out

    def _new_staged(self, batch: Dict[str, torch.Tensor]) -> _Staged:
        """Static device buffers of one batch shape: the mel on its own, everything else as views of one packed allocation."""
        offs, off = {}, 0
        for k in PACKED_KEYS:
            offs[k] = off
            off = (off + batch[k].numel() * batch[k].element_size() + 255) & ~255
        pack_dev = torch.empty(max(off, 256), dtype=torch.uint8, device=self.device)
        pack_host = torch.empty(PACK_RING, pack_dev.numel(), dtype=torch.uint8).pin_memory()

        def views(base: torch.Tensor) -> Dict[str, torch.Tensor]:
            return {k: base[offs[k]:offs[k] + batch[k].numel() * batch[k].element_size()].view(batch[k].dtype).view(batch[k].shape)
                    for k in PACKED_KEYS}
        dev = views(pack_dev)
        dev["mel_specs"] = torch.empty(batch["mel_specs"].shape, dtype=batch["mel_specs"].dtype, device=self.device)
        return _Staged(dev=dev, pack_dev=pack_dev, pack_host=pack_host, pack_views=[views(pack_host[i]) for i in range(PACK_RING)],
                       pack_done=[None] * PACK_RING)

    def stage(self, batch: Dict[str, torch.Tensor], divisor: int = 1) -> Tuple[_Staged, Tuple[int, int, int, int, bool]]:
        """Host-side prologue: shape key, stabiliser scalars, async H2D into the static buffers.
        divisor = the accumulation divisor of this micro-batch (trainer.py:2284-2294)."""
        validate_batch(batch)
        batch = self._cap(batch)
        dur = batch["phoneme_durations"]
        B, P = dur.shape
        T = batch["mel_specs"].shape[1]
        on_host = not dur.is_cuda
        if on_host:
            dsum = dur.clamp(min=0).sum(dim=1)
            Tp = max(1, int(dsum.max()))
            max_d = int(dur.max()) if dur.numel() else 0
        else:   # device-resident batch (bench `value` leg): caller guarantees sum(d) == T
            Tp, max_d = T, 0
        Tp = max(Tp, 3)
        # SpecAugment adds kernels to the step: graphs with and without it are different graphs
        key = (B, P, T, Tp, self.engine.spec_spans is not None)
        st = self._staged.get(key)
        self._tick += 1
        self._shape_hits += st is not None
        if st is None:
            if len(self._staged) >= self.max_cached_shapes:
                victim = min(self._staged, key=lambda k: self._staged[k].last_used)
                del self._staged[victim]          # drops the graph and its private memory pool
            st = self._new_staged(batch)
            self._staged[key] = st
        st.last_used = self._tick
        nbytes = 0
        packed = on_host and all(not batch[k].is_cuda and batch[k].dtype == st.dev[k].dtype for k in PACKED_KEYS)
        if packed:
            i = st.pack_i % PACK_RING
            st.pack_i += 1
            if st.pack_done[i] is not None:
                st.pack_done[i].synchronize()     # the copy that last read this pinned image has finished (PACK_RING steps ago)
            for k, view in st.pack_views[i].items():
                view.copy_(batch[k])
            st.pack_dev.copy_(st.pack_host[i], non_blocking=True)
            if st.pack_dev.is_cuda:
                if st.pack_done[i] is None:
                    st.pack_done[i] = torch.cuda.Event()
                st.pack_done[i].record()
            nbytes += sum(v.numel() * v.element_size() for v in st.pack_views[i].values())
        mel_idle, self._mel_idle = self._mel_idle, False
        for k in BATCH_KEYS:
            src = batch[k]
            if packed and k in PACKED_KEYS:
                continue
            if src.data_ptr() != st.dev[k].data_ptr():
                if k == "mel_specs" and mel_idle and packed and self._copy_stream is not None:
                    with torch.cuda.stream(self._copy_stream):
                        st.dev[k].copy_(src, non_blocking=True)
                        if self._mel_copied is None:
                            self._mel_copied = torch.cuda.Event()
                        self._mel_copied.record()
                    torch.cuda.current_stream(self.device).wait_event(self._mel_copied)
                else:
                    st.dev[k].copy_(src, non_blocking=True)
                if not src.is_cuda:
                    nbytes += src.numel() * src.element_size()
        scale, clip = adaptive_stabilisation(T, max_d, self.opt.cfg.max_grad_norm)
        slot = self._scal_ring[self._scal_i % self._scal_ring.shape[0]]
        self._scal_i += 1
        slot[0] = scale / (self.world * max(1, divisor))
        slot[1] = clip
        slot[2] = float(self._seq)
        self._scal_dev.copy_(slot, non_blocking=True)
        self.h2d_bytes_last_step = nbytes + 12
        return st, key

    # ------------------------------------------------------------------------------------------
    def _fwd_bwd_parts(self, d: Dict[str, torch.Tensor], Tp: int, zero: bool, out: list):
        """Generator over the (at most two) parts of zero-grad + forward + losses + backward; out[0] = losses."""
        eng = self.engine
        if zero:                                  # optimizer.zero_grad() at the start of a window, trainer.py:2258-2259
            with eng._on("z"):                    # 198 MB memset (33 us): beside the forward, the backward is its first reader
                eng.zero_grad()
        outs, ctx = eng.forward(d["phoneme_indices"], d["mel_specs"], d["phoneme_durations"], d["pitches"],
                                d["energies"], d["stress_indices"], expanded_len=Tp)
        losses, g = eng.losses(outs, d["mel_specs"], d["phoneme_durations"], d["stop_token_targets"],
                               d["pitches"], d["energies"], d["mel_lengths"], d["phoneme_lengths"],
                               loss_scale=self.loss_scale)
        out.append(losses)
        with eng._on("l"):                        # early D2H of the losses, then of the sequence id that marks them valid
            self._loss_host[0:6].copy_(losses, non_blocking=True)
            self._loss_host[6:7].copy_(self._scal_dev[2:3], non_blocking=True)
        if zero:
            eng._join("z")                        # every backward stream forks from here: the gradients are zero before any of them adds
        yield from eng.backward_parts(ctx, g, self.split_layer if getattr(self, "_early_on", False) else None)

    EARLY_GRID = 32      # blocks of the overlapped all-reduce: small enough to share the SMs with the backward kernels

    def _early_reduce(self, split_layer: int) -> None:
        """engine.early_reduce_hook: runs on the comm side stream (inside graph A when graphs are on)."""
        if self._early_on:
            self.reducer.reduce(chunk_ranges=self._early_ranges[split_layer], grid=self.EARLY_GRID)

    def _fwd_bwd(self, d: Dict[str, torch.Tensor], Tp: int, zero: bool = True) -> torch.Tensor:
        out: list = []
        for _ in self._fwd_bwd_parts(d, Tp, zero, out):
            pass
        return out[0]

    def _run_fwd_bwd(self, st: _Staged, key, zero: bool = True, early: bool = False) -> torch.Tensor:
        Tp = key[3]
        self._early_on = bool(early) and self.split_layer is not None
        v = int(zero) + 2 * int(self._early_on)
        if not self.use_graphs:
            n0 = launch_count()
            losses = self._fwd_bwd(st.dev, Tp, zero)
            st.launches = launch_count() - n0
            return losses
        if st.graph[v] is None:
            if st.warm[v] < 1 or not self._capture_worth():   # eager: warm-up (geometry tables, func attrs) or shapes that do not recur
                st.warm[v] += 1
                n0 = launch_count()
                losses = self._fwd_bwd(st.dev, Tp, zero)
                st.launches = launch_count() - n0
                return losses
            torch.cuda.synchronize(self.device)
            out: list = []
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in self._fwd_bwd_parts(st.dev, Tp, zero, out):
                    pass
            st.graph[v] = g
            st.graph_losses[v] = out[0]
        st.graph[v].replay()
        st.losses = st.graph_losses[v]
        return st.losses

    CAPTURE_PROBE = 32       # steps during which every recurring shape is captured
    CAPTURE_MIN_HIT = 0.5    # afterwards: fraction of steps whose shape must already be cached for capturing to go on

    def _capture_worth(self) -> bool:
        """Graph capture pays only if shapes come back: see the comment at _shape_hits."""
        return self._tick <= self.CAPTURE_PROBE or self._shape_hits >= self.CAPTURE_MIN_HIT * self._tick

    def _run_optimizer(self) -> None:
        fused = self.reducer is not None and self.world > 1
        # data parallel: the clip every replica applies is the minimum over the ranks' adaptive clips (written by the
        # collective kernel); single process: this batch's own
        clip = self.reducer.clip_global if fused else self.clip
        if not self.use_graphs or (self._opt_graph is None and self._opt_warm < 1):
            self._opt_warm += 1
            n0 = launch_count()
            self.opt.step(clip_override=clip, sq_chunk_ready=fused)
            self._opt_launches = launch_count() - n0
            return
        if self._opt_graph is None:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.opt.step(clip_override=clip, sq_chunk_ready=fused)
            self._opt_graph = g
        self._opt_graph.replay()

    # ------------------------------------------------------------------------------------------
    def micro_step(self, batch: Dict[str, torch.Tensor], first: bool = True, last: bool = True,
                   divisor: int = 1) -> torch.Tensor:
        """One micro-batch of a gradient-accumulation window (reference trainer.py:2258-2294, 2341-2343):
        `first` zeroes the gradient buffer, the loss is scaled by adaptive_scale / divisor, and `last` closes the
        window: one all-reduce (world > 1), pre-clip / clip (with THIS micro-batch's adaptive clip, as the
        reference does), AdamW, scheduler, EMA, projection.  Returns the device tensor
        losses[6] = (total, mel, duration, stop, pitch, energy), un-scaled."""
        st, key = self.stage(batch, divisor)
        if last:
            self.opt.set_lrs(self.sched.lrs())
        losses = self._run_fwd_bwd(st, key, zero=first, early=(last and self.world > 1))
        self.launches_last_step = st.launches
        if last:
            if self.world > 1:
                # what the overlapped launch did not cover (everything, without the overlap) + the global clip
                self.reducer.reduce(clip_local=self.clip,
                                    chunk_ranges=self._late_ranges if self.split_layer is not None else None)
            self._run_optimizer()
            self.sched.advance()
            self.launches_last_step += self._opt_launches + (1 if self.world > 1 else 0)
        return losses

    @torch.no_grad()
    def new_val_metrics(self) -> torch.Tensor:
        """Zeroed device accumulator for eval_losses(..., metrics=acc): one per validation epoch."""
        from . import ops
        return ops.zero_(torch.empty(ops.val_metrics_acc_floats(), dtype=torch.float32, device=self.device))

    @staticmethod
    def read_val_metrics(acc: torch.Tensor) -> Dict[str, Optional[float]]:
        """The epoch's ONE device read (trainer.py:1933-1934: None when no batch contributed)."""
        a = acc[:4].cpu().tolist()
        return {"val_spectral_convergence": a[0] / a[1] if a[1] > 0 else None,
                "val_f0_rmse": a[2] / a[3] if a[3] > 0 else None}

    @torch.no_grad()
    def eval_losses(self, batch: Dict[str, torch.Tensor], use_ema: bool = True,
                    metrics: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Validation forward (reference trainer.py:1771-1985, validate_epoch): the model in eval() mode — dropout and
        stochastic depth off, SpecAugment off — evaluated with the EMA weights (`use_ema`, trainer.py:1790-1806), returns
        the un-scaled losses[6] of the batch; with `metrics` (new_val_metrics()) the batch's spectral convergence and F0 RMSE
        are folded into that device accumulator.  Gradients, optimizer state and the RNG step are left untouched.  Eager
        (validation runs once per epoch); the bf16 operand copy of the EMA weights is a scratch buffer."""
        eng, st = self.engine, self.engine.store
        batch = self._cap(batch)
        dev = {k: batch[k].to(self.device, non_blocking=True) for k in BATCH_KEYS}
        dur = batch["phoneme_durations"]
        Tp = max(3, int(dur.clamp(min=0).sum(dim=1).max()))
        saved = (st.params, st.shadow, eng.training, eng.spec_spans)
        try:
            if use_ema and st.ema is not None:
                if getattr(self, "_ema_shadow", None) is None:
                    self._ema_shadow = torch.empty_like(st.shadow)
                from . import ops
                ops.cast_bf16(st.ema, self._ema_shadow)
                st.params, st.shadow = st.ema, self._ema_shadow
            eng.training, eng.spec_spans = False, None
            outs, _ = eng.forward(dev["phoneme_indices"], dev["mel_specs"], dev["phoneme_durations"], dev["pitches"],
                                  dev["energies"], dev["stress_indices"], expanded_len=Tp)
            losses, _ = eng.losses(outs, dev["mel_specs"], dev["phoneme_durations"], dev["stop_token_targets"],
                                   dev["pitches"], dev["energies"], dev["mel_lengths"], dev["phoneme_lengths"])
            if metrics is not None:     # spectral convergence / F0 RMSE accumulated on the device (trainer.py:1868-1916)
                from . import ops
                ops.val_metrics(outs[0], dev["mel_specs"], outs[3], dev["pitches"], dev["mel_lengths"], metrics)
            return losses.clone()
        finally:
            st.params, st.shadow, eng.training, eng.spec_spans = saved

    def train_step(self, batch: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Full optimizer step on one micro-batch (gradient_accumulation_steps = 1)."""
        return self.micro_step(batch, True, True, 1)

    def train_step_host(self, batch: Dict[str, torch.Tensor], timeout_s: float = 120.0) -> List[float]:
        """train_step() for callers that log the losses of EVERY step (the reference trainer's progress bar,
        trainer.py:2560-2600): returns the six un-scaled losses as Python floats as soon as the device has produced them —
        right after the forward pass — while the backward pass and the optimizer of the step are still running; the host
        stages the next batch in that time.  Stream order is unchanged: the next step starts after this one has finished."""
        self._seq = self._seq % 1000000 + 1
        want = float(self._seq)
        self.micro_step(batch, True, True, 1)
        flag = self._loss_host.numpy()
        t0 = time.monotonic()
        spins = 0
        while flag[6] != want:                    # pinned memory: the copy engine's writes are visible to host loads
            spins += 1
            if (spins & 0xFFFF) == 0 and time.monotonic() - t0 > timeout_s:
                torch.cuda.synchronize(self.device)
                if flag[6] != want:
                    raise RuntimeError("train_step_host: the losses of the step never arrived (sequence id %r, want %r)"
                                       % (float(flag[6]), want))
        self._mel_idle = True                     # this step is past the loss kernel, the last reader of the mel batch
        return [float(v) for v in flag[:6]]

    def train_window(self, batches: List[Dict[str, torch.Tensor]]) -> List[torch.Tensor]:
        """One optimizer step over an accumulation window of len(batches) micro-batches; the divisor is the
        window length (= min(G, remaining) of trainer.py:3345-3362 when the caller cuts the windows)."""
        n = len(batches)
        out = []
        for i, b in enumerate(batches):
            out.append(self.micro_step(b, first=(i == 0), last=(i == n - 1), divisor=n).clone())
        return out
