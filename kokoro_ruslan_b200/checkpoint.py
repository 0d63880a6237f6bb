"""Checkpoint pieces of the training path in the reference's on-disk format
(src/kokoro/training/trainer.py:1994-2031, ``save_checkpoint_with_scaler``; resume rules
training/checkpoint_manager.py:360-525).

* ``optimizer_state_dict`` / ``load_optimizer_state_dict`` — the fused optimizer's flat Adam moment buffers presented
  as (and restored from) a ``torch.optim.AdamW.state_dict()``: ten param groups in the reference's order
  (``_setup_optimizer``, trainer.py:446-689; SURVEY A13'), parameters numbered consecutively group by group in
  ``named_parameters()`` order, per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` in the reference's tensor shapes
  (conv weights permuted back from the tap-major layout the kernels use).  A checkpoint whose group layout differs is
  not an error: like the reference (checkpoint_manager.py:478-507) the moments then restart from zero.
* ``AsyncCheckpointWriter`` — device → pinned-host snapshot on a side stream, ``torch.save`` on a worker thread: the
  training stream waits for the D2H copies of the ~0.8 GB state (weights, EMA, two moment buffers; tens of milliseconds
  over PCIe 5), not for serialisation or the file system.

Pure tensor plumbing: no kernels, works on whatever device the ParamStore lives on.
"""
from __future__ import annotations

import os
import threading
from typing import Callable, Dict, List, Optional

import torch

from .optim import CTRL_FIELDS, group_hparams, group_of

GROUP_NAMES = ["encoder", "encoder_ffn_decay", "decoder_other_no_decay", "decoder_other_decay", "decoder_attn_decay",
               "decoder_attn_no_decay", "decoder_ffn_decay", "decoder_ffn_no_decay", "variance_embed", "stop_head"]
_STEP = CTRL_FIELDS.index("step")


def group_layout(names: List[str]) -> List[List[str]]:
    """Parameter names per AdamW group, each in named_parameters() order (how the reference fills its groups)."""
    groups: List[List[str]] = [[] for _ in GROUP_NAMES]
    for n in names:
        groups[group_of(n)].append(n)
    return groups


def optimizer_state_dict(opt, lrs: Optional[List[float]] = None, step: Optional[int] = None, on_device: bool = False) -> Dict:
    """``torch.optim.AdamW.state_dict()``-shaped snapshot of a FusedAdamW: CPU tensors, or — on_device, for save_sharded,
    which copies only a rank's share — views of the device moment buffers."""
    st, cfg = opt.store, opt.cfg
    if step is None:
        step = int(opt.ctrl.cpu()[_STEP])
    if lrs is None:
        lrs = [float(x) for x in opt.g_lr.cpu().tolist()]
    hp = group_hparams(cfg)
    take = (lambda t: t.detach()) if on_device else (lambda t: t.detach().cpu().contiguous().clone())
    state: Dict[int, Dict[str, torch.Tensor]] = {}
    param_groups = []
    idx = 0
    for g, names in enumerate(group_layout(st.order)):
        ids = []
        for n in names:
            if step > 0:                              # torch creates a parameter's state at its first step
                state[idx] = {"step": torch.tensor(float(step)),
                              "exp_avg": take(st.ref_view(st.exp_avg, n)), "exp_avg_sq": take(st.ref_view(st.exp_avg_sq, n))}
            ids.append(idx)
            idx += 1
        param_groups.append({"lr": float(lrs[g]), "betas": tuple(cfg.adam_betas), "eps": cfg.adam_eps,
                             "weight_decay": float(hp[g][1]), "amsgrad": False, "maximize": False, "foreach": None,
                             "capturable": False, "differentiable": False, "fused": True,
                             "decoupled_weight_decay": True, "name": GROUP_NAMES[g],
                             "initial_lr": cfg.learning_rate * hp[g][0], "params": ids})
    return {"state": state, "param_groups": param_groups}


def load_optimizer_state_dict(opt, sd: Optional[Dict], log: Callable[[str], None] = lambda s: None) -> bool:
    """Restores the Adam moments and the step counter.  Returns False (moments zeroed, counter kept at the caller's
    value) when the checkpoint's group layout does not match — the reference's behaviour for a changed group count."""
    st = opt.store
    layout = group_layout(st.order)

    def restart(why: str) -> bool:
        log(f"optimizer state not loaded ({why}): Adam moments restart from zero")
        st.exp_avg.zero_()
        st.exp_avg_sq.zero_()
        return False

    if not sd or "param_groups" not in sd or "state" not in sd:
        return restart("no optimizer_state_dict")
    pgs = sd["param_groups"]
    if len(pgs) != len(layout) or any(len(pg["params"]) != len(names) for pg, names in zip(pgs, layout)):
        return restart(f"{len(pgs)} param groups of sizes {[len(pg['params']) for pg in pgs]} in the checkpoint, "
                       f"{[len(n) for n in layout]} here")
    state = sd["state"]
    step = 0
    with torch.no_grad():
        for pg, names in zip(pgs, layout):
            for pid, n in zip(pg["params"], names):
                s = state.get(pid, state.get(str(pid)))
                if s is None:                                      # a parameter that never received a gradient
                    st.ref_view(st.exp_avg, n).zero_()
                    st.ref_view(st.exp_avg_sq, n).zero_()
                    continue
                want = tuple(st.ref_view(st.exp_avg, n).shape)
                if tuple(s["exp_avg"].shape) != want:
                    return restart(f"shape of {n}: {tuple(s['exp_avg'].shape)} vs {want}")
                st.ref_view(st.exp_avg, n).copy_(s["exp_avg"].to(st.device, torch.float32))
                st.ref_view(st.exp_avg_sq, n).copy_(s["exp_avg_sq"].to(st.device, torch.float32))
                step = max(step, int(float(s["step"])))
        ctrl = opt.ctrl.cpu()
        ctrl[_STEP] = step
        opt.ctrl.copy_(ctrl)
    return True


def build_model_metadata(model_cfg, run_cfg=None) -> Dict:
    """``model_metadata`` of the reference's checkpoints (training/checkpoint_manager.py:178-241, schema 2): the
    architecture record its resume path demands (`checkpoint_manager.py:360-420`) and the inference controls.  Built from
    the ModelConfig the engine was constructed with (the authoritative record here — the reference re-derives the FFN
    widths from module attributes for the same reason) and the run's dropout settings."""
    g = lambda name, default: getattr(run_cfg, name, default) if run_cfg is not None else default      # noqa: E731
    return {
        "schema_version": 2,
        "architecture": {
            "mel_dim": int(model_cfg.mel_dim), "hidden_dim": int(model_cfg.hidden_dim),
            "n_encoder_layers": int(model_cfg.n_encoder_layers), "n_decoder_layers": int(model_cfg.n_decoder_layers),
            "n_heads": int(model_cfg.n_heads), "encoder_ff_dim": int(model_cfg.encoder_ff_dim),
            "decoder_ff_dim": int(model_cfg.decoder_ff_dim), "encoder_dropout": float(g("encoder_dropout", 0.15)),
            "max_decoder_seq_len": int(model_cfg.max_decoder_seq_len), "use_variance_predictor": True,
            "variance_filter_size": int(model_cfg.variance_filter_size),
            "variance_kernel_size": int(model_cfg.variance_kernel_size),
            "variance_dropout": float(g("variance_dropout", 0.1)), "n_variance_bins": int(model_cfg.n_variance_bins),
            "pitch_min": 0.0, "pitch_max": 1.0, "energy_min": 0.0, "energy_max": 1.0,
            "use_stochastic_depth": float(g("stochastic_depth_rate", 0.1)) > 0.0,
            "stochastic_depth_rate": float(g("stochastic_depth_rate", 0.1)), "qk_norm": bool(model_cfg.qk_norm),
            "ffn_output_norm": bool(model_cfg.ffn_output_norm), "vocab_size": int(model_cfg.vocab_size),
        },
        "inference_controls": {"max_len": 1200, "stop_threshold": 0.45, "min_len_ratio": 0.7, "min_len_floor": 12},
    }


def check_model_metadata(meta: Optional[Dict], model_cfg) -> List[str]:
    """Differences between a checkpoint's architecture record and this engine (empty = compatible).  The reference
    refuses to resume on a mismatch (checkpoint_manager.py:380-420); a checkpoint without the record is let through —
    the strict state-dict load that follows is the real gate."""
    if not meta or "architecture" not in meta:
        return []
    arch, bad = meta["architecture"], []
    for key in ("mel_dim", "hidden_dim", "n_encoder_layers", "n_decoder_layers", "n_heads", "encoder_ff_dim",
                "decoder_ff_dim", "variance_filter_size", "variance_kernel_size", "n_variance_bins", "vocab_size"):
        if key in arch and int(arch[key]) != int(getattr(model_cfg, key)):
            bad.append(f"{key}: checkpoint {arch[key]} vs model {getattr(model_cfg, key)}")
    for key, want in (("pitch_min", 0.0), ("pitch_max", 1.0), ("energy_min", 0.0), ("energy_max", 1.0)):
        if key in arch and float(arch[key]) != want:
            bad.append(f"{key}: checkpoint {arch[key]} vs {want} (targets are normalised to [0, 1])")
    if arch.get("use_variance_predictor") is False:
        bad.append("use_variance_predictor: False in the checkpoint")
    return bad


class AsyncCheckpointWriter:
    """``writer.save(path, tensors_and_scalars)``: snapshots every tensor of the (nested) dict into host memory — pinned
    and through a side stream when it lives on a CUDA device — and hands the file write to a worker thread.
    ``wait()`` blocks until every pending file is on disk (call it before reading a checkpoint back, and at exit)."""

    def __init__(self):
        self._threads: List[threading.Thread] = []
        self._errors: List[BaseException] = []
        self._stream = None

    def _snapshot(self, obj):
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                host = torch.empty(obj.shape, dtype=obj.dtype, pin_memory=True)
                host.copy_(obj.detach(), non_blocking=True)
                return host
            return obj.detach().clone()
        if isinstance(obj, dict):
            return {k: self._snapshot(v) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._snapshot(v) for v in obj)
        return obj

    def save(self, path: str, payload: Dict) -> None:
        event = None
        if torch.cuda.is_available() and any(t.is_cuda for t in _tensors(payload)):
            if self._stream is None:
                self._stream = torch.cuda.Stream()
            self._stream.wait_stream(torch.cuda.current_stream())      # the snapshot sees the state as of this call
            with torch.cuda.stream(self._stream):
                snap = self._snapshot(payload)
                event = torch.cuda.Event()
                event.record(self._stream)
            # the training stream must not overwrite the buffers before the copies have read them
            torch.cuda.current_stream().wait_stream(self._stream)
        else:
            snap = self._snapshot(payload)

        def work():
            try:
                if event is not None:
                    event.synchronize()
                tmp = path + ".tmp"
                torch.save(snap, tmp)
                os.replace(tmp, path)                                   # readers never see a half-written file
            except BaseException as exc:                               # surfaced by wait()
                self._errors.append(exc)

        t = threading.Thread(target=work, name="kokoro-ckpt-writer", daemon=False)
        t.start()
        self._threads.append(t)

    def wait(self) -> None:
        for t in self._threads:
            t.join()
        self._threads = []
        if self._errors:
            err, self._errors = self._errors[0], []
            raise RuntimeError(f"asynchronous checkpoint write failed: {err}") from err


def _tensors(obj):
    if isinstance(obj, torch.Tensor):
        yield obj
    elif isinstance(obj, dict):
        for v in obj.values():
            yield from _tensors(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from _tensors(v)


# ----------------------------------------------------------------------------------------------------------------------
# Sharded checkpoints.  The data-parallel replicas hold IDENTICAL state (weights, EMA, Adam moments:
# tests/test_dp_gpu.py), so no gather is needed: rank r copies only ITS share of the tensors off the device and writes
# `<path>.shard<r>of<N>`; rank 0 also writes `<path>` itself with every non-tensor entry and the table of which shard holds
# which tensor.  `load_sharded(path)` rebuilds exactly the dict a single-file `torch.save(payload, path)` would have held,
# so `cli.resume`, the reference's `torch.load(...)["model_state_dict"]` consumers and the EMA export read it unchanged
# after one merge.  Device-to-host volume and file-system time per rank drop by the world size.
# ----------------------------------------------------------------------------------------------------------------------
_SHARD_KEY = "__kokoro_b200_sharded__"


def _tensor_paths(obj, prefix=()):
    if isinstance(obj, torch.Tensor):
        yield prefix, obj
    elif isinstance(obj, dict):
        for k, v in obj.items():
            yield from _tensor_paths(v, prefix + (k,))
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            yield from _tensor_paths(v, prefix + (i,))


def shard_assignment(payload: Dict, world: int) -> Dict[tuple, int]:
    """tensor path -> rank: largest tensors first, each to the currently lightest rank (deterministic on every rank)."""
    items = sorted(_tensor_paths(payload), key=lambda kv: (-kv[1].numel() * kv[1].element_size(), repr(kv[0])))
    load = [0] * world
    out = {}
    for path, t in items:
        r = min(range(world), key=lambda i: (load[i], i))
        out[path] = r
        load[r] += t.numel() * t.element_size()
    return out


def _strip(obj, prefix, table):
    """payload with every tensor replaced by a (marker, rank) placeholder"""
    if isinstance(obj, torch.Tensor):
        return (_SHARD_KEY, table[prefix])
    if isinstance(obj, dict):
        return {k: _strip(v, prefix + (k,), table) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_strip(v, prefix + (i,), table) for i, v in enumerate(obj))
    return obj


def shard_path(path: str, rank: int, world: int) -> str:
    return f"{path}.shard{rank}of{world}"


def save_sharded(path: str, payload: Dict, rank: int, world: int, writer: Optional["AsyncCheckpointWriter"] = None) -> List[str]:
    """Called by EVERY rank with the same (replicated) payload; tensors may live on the device.  Returns the files this rank
    wrote (or queued on `writer`)."""
    table = shard_assignment(payload, world)
    mine = {repr(p): t for p, t in _tensor_paths(payload) if table[p] == rank}
    files = [(shard_path(path, rank, world), {"rank": rank, "world": world, "tensors": mine})]
    if rank == 0:
        files.append((path, {_SHARD_KEY: {"world": world, "version": 1}, "skeleton": _strip(payload, (), table)}))
    for f, obj in files:
        if writer is not None:
            writer.save(f, obj)                                        # snapshots this rank's tensors only
        else:
            if "tensors" in obj:
                obj = dict(obj, tensors={k: v.detach().cpu() for k, v in obj["tensors"].items()})
            tmp = f + ".tmp"
            torch.save(obj, tmp)
            os.replace(tmp, f)
    return [f for f, _ in files]


def is_sharded(obj) -> bool:
    return isinstance(obj, dict) and _SHARD_KEY in obj


def load_sharded(path: str, map_location="cpu") -> Dict:
    """The merged payload of a sharded checkpoint; a plain single-file checkpoint is returned as it is."""
    head = torch.load(path, map_location=map_location, weights_only=False)
    if not is_sharded(head):
        return head
    world = int(head[_SHARD_KEY]["world"])
    shards = []
    for r in range(world):
        f = shard_path(path, r, world)
        if not os.path.exists(f):
            raise FileNotFoundError(f"sharded checkpoint {path}: shard {r} of {world} is missing ({f})")
        s = torch.load(f, map_location=map_location, weights_only=False)
        if s.get("rank") != r or s.get("world") != world:
            raise RuntimeError(f"{f}: written as shard {s.get('rank')} of {s.get('world')}, expected {r} of {world}")
        shards.append(s["tensors"])

    def fill(obj, prefix):
        if isinstance(obj, tuple) and len(obj) == 2 and obj[0] == _SHARD_KEY:
            try:
                return shards[obj[1]][repr(prefix)]
            except KeyError:
                raise RuntimeError(f"sharded checkpoint {path}: tensor {prefix} is not in shard {obj[1]}") from None
        if isinstance(obj, dict):
            return {k: fill(v, prefix + (k,)) for k, v in obj.items()}
        if isinstance(obj, (list, tuple)):
            return type(obj)(fill(v, prefix + (i,)) for i, v in enumerate(obj))
        return obj
    return fill(head["skeleton"], ())


# ----------------------------------------------------------------------------------------------------------------------
# Legacy checkpoints: the reference's strict-resume rules (training/checkpoint_manager.py:360-525) restated on state dicts.
# ----------------------------------------------------------------------------------------------------------------------
_NEW_VARIANCE_PREFIX = "duration_adaptor.variance_adaptor."       # sub-modules added in a later architecture revision
_FFN_OUTPUT_NORM_SUFFIX = ".ff.output_norm.weight"                # enabled on a checkpoint trained without it
_ALIBI_SUFFIX = ".alibi_slopes"                                   # buffers of the ALiBi decoder, gone with RoPE


def extract_model_state_dict(ck) -> Dict[str, torch.Tensor]:
    """`model_state_dict`, else `model`, else a raw state dict (checkpoint_manager.py:398-410)."""
    if not isinstance(ck, dict):
        raise RuntimeError("Checkpoint payload is not a dictionary.")
    if "model_state_dict" in ck:
        return ck["model_state_dict"]
    if "model" in ck:
        return ck["model"]
    if ck and all(isinstance(v, torch.Tensor) for v in ck.values()):
        return ck
    raise RuntimeError("Checkpoint does not contain a recognized model state dictionary. "
                       "Expected key 'model_state_dict' or 'model'.")


def migrate_model_state_dict(saved: Dict[str, torch.Tensor], current: Dict[str, torch.Tensor],
                             log: Callable[[str], None] = lambda s: None) -> Dict[str, torch.Tensor]:
    """The state dict to load strictly into the current model, after the reference's known migrations
    (checkpoint_manager.py:412-489): keys missing from `saved` are tolerated only under
    `duration_adaptor.variance_adaptor.` or as `*.ff.output_norm.weight` — they keep the current (freshly initialised)
    values; keys `saved` has in excess are tolerated only as `*.alibi_slopes` and are discarded.  Anything else, or a shape
    mismatch, is the reference's "architecture/state mismatch" error."""
    missing = [k for k in current if k not in saved]
    unexpected = [k for k in saved if k not in current]
    shapes = [k for k in current if k in saved and tuple(saved[k].shape) != tuple(current[k].shape)]
    ok_missing = all(k.startswith(_NEW_VARIANCE_PREFIX) or k.endswith(_FFN_OUTPUT_NORM_SUFFIX) for k in missing)
    ok_unexpected = all(k.endswith(_ALIBI_SUFFIX) for k in unexpected)
    if shapes or not (ok_missing and ok_unexpected):
        raise RuntimeError("Strict checkpoint model load failed due to architecture/state mismatch. "
                           f"Missing: {[k for k in missing if not (k.startswith(_NEW_VARIANCE_PREFIX) or k.endswith(_FFN_OUTPUT_NORM_SUFFIX))]}, "
                           f"unexpected: {[k for k in unexpected if not k.endswith(_ALIBI_SUFFIX)]}, shape mismatches: {shapes}")
    if missing:
        log(f"Checkpoint is missing {len(missing)} key(s) from expected architecture migrations (variance_adaptor / "
            "ffn_output_norm). Loading shared weights and initialising missing keys from scratch.")
    if unexpected:
        log(f"Checkpoint contains {len(unexpected)} legacy ALiBi positional encoding buffer(s) (alibi_slopes) that are not "
            "used by the current RoPE-based model.  These will be discarded.")
    return {k: (saved[k] if k in saved else current[k].detach().clone()) for k in current}


def check_resume_fields(ck: Dict, training: bool = True) -> None:
    """Fields a training resume needs (checkpoint_manager.py:491-511): optimizer + scheduler state, epoch, loss."""
    if training and ("optimizer_state_dict" not in ck or "scheduler_state_dict" not in ck):
        raise RuntimeError("Checkpoint is missing optimizer/scheduler state required for training resume.")
    if "epoch" not in ck or "loss" not in ck:
        raise RuntimeError("Checkpoint is missing required 'epoch' or 'loss' fields.")
