"""``kokoro-train`` — the reference's training CLI (src/kokoro/cli/cli.py:33-290, console script
``kokoro-train = kokoro.cli.training:main``) driving the B200 training step.

Same flags, short options, defaults and ``dest`` names as the reference parser; the loop restates the
epoch structure of ``KokoroTrainer.train`` around ``TrainStep``:

  * batches from ``DynamicFrameBatchSampler`` (or ``LengthBasedBatchSampler`` with ``--no-dynamic-batching``),
    rebuilt every epoch, rank-sliced by ``DistributedBatchSampler`` under torchrun;
  * gradient-accumulation windows of ``gradient_accumulation_steps`` micro-batches with the exact divisor for
    the tail window (trainer.py:2258-2294, 3345-3362);
  * SpecAugment on the decoder memory from epoch index >= 1 (trainer.py:2042-2055);
  * epoch metrics = mean of the un-scaled micro-batch losses (trainer.py:2739-2748), read from the device ONCE
    per epoch;
  * validation on the EMA weights every ``--validation-interval`` epochs, early stopping with
    ``--early-stopping-patience`` and the 0.001 improvement margin (trainer.py:2943-2956);
  * ``checkpoint_epoch_{N}.pth`` every ``--save-every`` epochs and on improvement, with the reference's key
    names for the parts this path owns (model / EMA state dicts, counters, losses).

The corpus side (phoneme front-end, MFA durations, feature cache) is the reference's ``RuslanDataset`` and is
imported from the reference package when it is installed; ``--synthetic N`` (not a reference flag) trains on N
generated utterances instead, which is what the smoke test uses.  Flags that only make sense for the reference's
runtime (AMP profiling, fused-AdamW selection, MPS) are accepted and ignored with a note.
"""
from __future__ import annotations

import argparse
import os
import random
import sys
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch

def build_parser() -> argparse.ArgumentParser:
    """The reference parser (cli/cli.py:33-222), option for option."""
    p = argparse.ArgumentParser(description="Kokoro Language Model Training Script (B200 path)",
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--corpus", "-c", type=str, default="./ruslan_corpus")
    p.add_argument("--output", "-o", type=str, default="./kokoro_russian_model")
    p.add_argument("--resume", "-r", type=str, default=None)
    p.add_argument("--batch-size", "-b", type=int, default=8)
    p.add_argument("--epochs", "-e", type=int, default=None)
    p.add_argument("--learning-rate", "-lr", type=float, default=None)
    p.add_argument("--save-every", type=int, default=5)
    p.add_argument("--mfa-alignments", type=str, default=None)
    p.add_argument("--no-mfa", action="store_true")
    p.add_argument("--val-split", type=float, default=0.1)
    p.add_argument("--no-validation", action="store_true")
    p.add_argument("--early-stopping-patience", type=int, default=10)
    p.add_argument("--validation-interval", type=int, default=1)
    p.add_argument("--dynamic-batching", action="store_true", default=True)
    p.add_argument("--no-dynamic-batching", action="store_false", dest="dynamic_batching")
    p.add_argument("--max-frames", type=int, default=None)
    p.add_argument("--min-batch-size", type=int, default=4)
    p.add_argument("--max-batch-size", type=int, default=32)
    p.add_argument("--profile-amp", action="store_true")
    amp = p.add_mutually_exclusive_group()
    amp.add_argument("--enable-amp", action="store_true")
    amp.add_argument("--disable-amp", action="store_true")
    p.add_argument("--profile-amp-batches", type=int, default=10)
    p.add_argument("--fused-adamw", action="store_true")
    p.add_argument("--no-fused-adamw", action="store_true")
    p.add_argument("--try-fused-adamw-mps", action="store_true", default=True)
    p.add_argument("--verbose", "-v", action="store_true")
    p.add_argument("--no-memory-cache", action="store_false", dest="use_memory_cache")
    p.add_argument("--stop-threshold", dest="stop_threshold", type=float, default=0.1)
    # not in the reference: corpus-free runs
    p.add_argument("--synthetic", type=int, default=0, metavar="N",
                   help="train on N generated utterances instead of a corpus (smoke / benchmarking)")
    p.add_argument("--seed", type=int, default=42)
    return p


@dataclass
class RunConfig:
    """The subset of the reference TrainingConfig this loop consumes (training/config.py; values the CLI does not
    set keep the reference defaults)."""
    data_dir: str = "./ruslan_corpus"
    output_dir: str = "./kokoro_russian_model"
    batch_size: int = 8
    num_epochs: int = 30
    learning_rate: float = 5.0e-5
    gradient_accumulation_steps: int = 2
    save_every: int = 5
    use_dynamic_batching: bool = True
    max_frames_per_batch: int = 30000
    min_batch_size: int = 4
    max_batch_size: int = 32
    validation_split: float = 0.1
    validation_interval: int = 1
    early_stopping_patience: int = 10
    early_stopping_min_delta: float = 0.001
    spec_augment_start_epoch: int = 1
    encoder_dropout: float = 0.15
    decoder_dropout: float = 0.20
    decoder_input_dropout: float = 0.15
    variance_dropout: float = 0.1
    stochastic_depth_rate: float = 0.1
    resume_checkpoint: Optional[str] = None
    verbose: bool = False
    seed: int = 42
    ema_decay: Optional[float] = None         # None: computed from the steps per epoch (config.py:85-86)
    ema_half_life_epochs: float = 1.0
    # spectral convergence / F0 RMSE of the validation epoch as device reductions (reference trainer.py:1868-1916 always
    # computes them, with per-utterance host syncs); off by default until kr_val_metrics has had its first hardware run
    val_metrics: bool = False
    async_checkpoints: bool = False           # torch.save on a writer thread (checkpoint.AsyncCheckpointWriter)
    # data-parallel runs: every rank writes 1 / world of the tensors (checkpoint.save_sharded; the replicas are identical, so
    # nothing is gathered); `resume` and checkpoint.load_sharded read both forms.  Off = the reference's single file.
    sharded_checkpoints: bool = False


class TrainingConfig(RunConfig):
    """What checkpoints pickle as their ``config`` entry.  The reference stores a pickled
    ``kokoro.training.config.TrainingConfig`` INSTANCE there (trainer.py:1994-2031, registered as a safe global at
    trainer.py:47) and its resume path reads attributes off it, so the class advertises that import path: a checkpoint
    written here unpickles, in a reference environment, into the reference's own dataclass (pickle restores the attribute
    dict onto it — same field names), and in an environment with the ``kokoro`` shim of this repository (shim/kokoro) into
    this class."""


TrainingConfig.__module__ = "kokoro.training.config"
TrainingConfig.__qualname__ = "TrainingConfig"


def _picklable_config(cfg: RunConfig):
    """cfg as an instance of whatever ``kokoro.training.config.TrainingConfig`` is importable here — the shim's class
    (= TrainingConfig above) or the reference's own dataclass when the reference package is installed.  The in-tree shim
    directory is put on sys.path if nothing named ``kokoro`` is installed; a plain dict only if even that fails."""
    import importlib
    try:
        mod = importlib.import_module("kokoro.training.config")
    except ImportError:
        shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shim")
        if not os.path.isdir(os.path.join(shim, "kokoro")):
            return dict(cfg.__dict__)
        sys.path.append(shim)
        try:
            mod = importlib.import_module("kokoro.training.config")
        except ImportError:
            return dict(cfg.__dict__)
    out = mod.TrainingConfig()
    for k, v in cfg.__dict__.items():
        setattr(out, k, v)
    return out


def create_config_from_args(args) -> RunConfig:
    """cli/cli.py:225-290."""
    cfg = TrainingConfig(data_dir=args.corpus, output_dir=args.output, batch_size=args.batch_size, save_every=args.save_every,
                    use_dynamic_batching=args.dynamic_batching,
                    max_frames_per_batch=args.max_frames if args.max_frames is not None else 30000,
                    min_batch_size=args.min_batch_size, max_batch_size=args.max_batch_size,
                    validation_split=0.0 if args.no_validation else args.val_split,
                    validation_interval=args.validation_interval, early_stopping_patience=args.early_stopping_patience,
                    resume_checkpoint=args.resume, verbose=args.verbose, seed=args.seed)
    if args.learning_rate is not None:
        cfg.learning_rate = args.learning_rate
    if args.epochs is not None:
        cfg.num_epochs = args.epochs
    return cfg


# ----------------------------------------------------------------------------------------------
# datasets
# ----------------------------------------------------------------------------------------------
class SyntheticDataset:
    """N generated utterances with the item layout of RuslanDataset.__getitem__ (data/dataset.py:849-862)."""

    def __init__(self, n: int, vocab_size: int = 59, n_mels: int = 80, seed: int = 0, min_frames: int = 200,
                 max_frames: int = 800):
        from .data import build_stop_token_targets
        g = torch.Generator().manual_seed(seed)
        self.items: List[Dict] = []
        self.samples: List[Dict] = []
        for i in range(n):
            T = int(torch.randint(min_frames, max_frames + 1, (1,), generator=g))
            P = max(8, T // 7)
            dur = torch.full((P,), T // P, dtype=torch.long)
            dur[: T - (T // P) * P] += 1
            self.items.append({
                "mel_spec": torch.randn(n_mels, T, generator=g) * 2.0 - 5.0,
                "phoneme_indices": torch.randint(1, vocab_size, (P,), generator=g),
                "stress_indices": torch.randint(0, 3, (P,), generator=g),
                "phoneme_durations": dur, "stop_token_targets": build_stop_token_targets(T),
                "pitch": torch.rand(T, generator=g), "energy": torch.rand(T, generator=g),
                "mel_length": T, "phoneme_length": P, "text": f"synthetic {i}", "audio_file": f"synthetic_{i}.wav"})
            self.samples.append({"audio_length": T})       # mel frames, as the reference indexes its corpus (dataset.py:311-338)
        self.vocab_size = vocab_size

    def __len__(self) -> int:
        return len(self.items)

    def __getitem__(self, i: int) -> Dict:
        return self.items[i]


class _Subset:
    def __init__(self, ds, indices: Sequence[int]):
        self.ds, self.indices = ds, list(indices)
        self.samples = [ds.samples[i] for i in self.indices]
        self.thread_safe = getattr(ds, "thread_safe", False)

    def __len__(self) -> int:
        return len(self.indices)

    def __getitem__(self, i: int):
        return self.ds[self.indices[i]]


def load_reference_dataset(cfg: RunConfig):
    """RuslanDataset of the installed reference package (corpus parsing, phonemes, MFA durations, feature cache are
    the reference's own code: out of this repository's scope)."""
    try:
        from kokoro.data.dataset import RuslanDataset                 # type: ignore
        from kokoro.training.config import TrainingConfig             # type: ignore
    except ImportError as exc:
        # no reference package: a corpus whose features the reference has already cached (<corpus>/.feature_cache/*.pt,
        # dataset.py:412-414,849-866) can still be trained on — the cache files hold everything the step consumes
        cache_dir = os.path.join(cfg.data_dir, ".feature_cache")
        if os.path.isdir(cache_dir):
            from .feature_cache import CachedFeatureDataset, FeatureCache
            from .params import ModelConfig
            ds = CachedFeatureDataset.scan(FeatureCache(cache_dir))
            if len(ds) > 0:
                top = max(int(ds.cache.load(s["audio_file"])["phoneme_indices"].max()) for s in ds.samples)
                ds.vocab_size = max(ModelConfig().vocab_size, top + 1)
                return ds
        raise RuntimeError("a corpus run needs the reference package (`kokoro`) on PYTHONPATH for RuslanDataset, or a "
                           "feature cache written by it under <corpus>/.feature_cache; use --synthetic N for a corpus-free run") from exc
    tc = TrainingConfig(data_dir=cfg.data_dir, output_dir=cfg.output_dir, batch_size=cfg.batch_size)
    ds = RuslanDataset(cfg.data_dir, tc)
    ds.vocab_size = len(ds.phoneme_processor.phoneme_to_id)
    return ds


def split_dataset(ds, val_split: float, seed: int):
    n = len(ds)
    if val_split <= 0.0 or n < 2:
        return ds, None
    idx = list(range(n))
    random.Random(seed).shuffle(idx)
    n_val = max(1, int(n * val_split))
    return _Subset(ds, idx[n_val:]), _Subset(ds, idx[:n_val])


# ----------------------------------------------------------------------------------------------
# epoch loop
# ----------------------------------------------------------------------------------------------
def accumulation_windows(n_batches: int, g: int) -> List[List[int]]:
    """Batch indices grouped into accumulation windows; the tail window is shorter and its divisor is its own
    length (_effective_accumulation_divisor, trainer.py:3345-3362)."""
    g = max(1, int(g))
    return [list(range(s, min(s + g, n_batches))) for s in range(0, n_batches, g)]


def make_sampler(ds, cfg: RunConfig, shuffle: bool = True):
    from .data import DynamicFrameBatchSampler, LengthBasedBatchSampler
    if cfg.use_dynamic_batching:
        return DynamicFrameBatchSampler(ds, max_frames=cfg.max_frames_per_batch, min_batch_size=cfg.min_batch_size,
                                        max_batch_size=cfg.max_batch_size, drop_last=False, shuffle=shuffle)
    return LengthBasedBatchSampler(ds, cfg.batch_size, drop_last=False, shuffle=shuffle)


def find_latest_checkpoint(output_dir: str) -> Optional[str]:
    """`--resume auto`: the checkpoint_epoch_{N}.pth with the largest N (reference checkpoint_manager semantics)."""
    import re
    best, best_n = None, -1
    if os.path.isdir(output_dir):
        for name in os.listdir(output_dir):
            m = re.fullmatch(r"checkpoint_epoch_(\d+)\.pth", name)
            if m and int(m.group(1)) > best_n:
                best, best_n = os.path.join(output_dir, name), int(m.group(1))
    return best


def resume(cfg: RunConfig, step, log: Callable[[str], None] = print) -> int:
    """Loads weights (+ EMA, scheduler counters) from cfg.resume_checkpoint ("auto" = latest in the output directory);
    returns the epoch index to continue with.  Adam moments are restored from `optimizer_state_dict` when its ten-group
    layout matches (ours, or a reference checkpoint written with the un-collapsed groups); otherwise they restart from
    zero, which is also what the reference does for a changed group count (checkpoint_manager.py:478-507)."""
    path = cfg.resume_checkpoint
    if not path:
        return 0
    if path == "auto":
        path = find_latest_checkpoint(cfg.output_dir)
        if path is None:
            log(f"--resume auto: no checkpoint in {cfg.output_dir}, starting from scratch")
            return 0
    from .checkpoint import load_sharded
    ck = load_sharded(path, map_location="cpu")       # a plain single-file checkpoint comes back as it is
    model_cfg = getattr(getattr(step, "engine", None), "cfg", None)
    if model_cfg is not None and hasattr(model_cfg, "hidden_dim"):
        from .checkpoint import check_model_metadata
        bad = check_model_metadata(ck.get("model_metadata"), model_cfg)
        if bad:
            raise RuntimeError(f"checkpoint {path} was written for a different architecture: " + "; ".join(bad))
    from .checkpoint import check_resume_fields, extract_model_state_dict, migrate_model_state_dict
    check_resume_fields(ck, training=hasattr(step, "opt"))
    # the reference's known migrations (checkpoint_manager.py:412-489): later-added variance-adaptor / ffn output-norm
    # weights keep their initial values, legacy ALiBi buffers are dropped; anything else is an architecture mismatch
    step.load_state_dict(migrate_model_state_dict(extract_model_state_dict(ck), step.state_dict(), log))
    ema = ck.get("ema_model_state_dict")
    st = step.store
    if ema is not None and getattr(st, "ema", None) is not None:
        live = st.params.clone()
        step.load_state_dict(migrate_model_state_dict(ema, step.state_dict()))   # same layout conversion + migrations ...
        st.ema.copy_(st.params)
        st.params.copy_(live)                     # ... then restore the live weights and their bf16 shadow
        st.refresh_shadow()
    if "scheduler_state_dict" in ck and hasattr(step.sched, "load_state_dict"):
        try:
            step.sched.load_state_dict(ck["scheduler_state_dict"])
        except (KeyError, TypeError):
            pass                                   # a reference-trainer checkpoint: its OneCycleLR state does not apply
    if hasattr(step, "opt"):
        from .checkpoint import load_optimizer_state_dict
        load_optimizer_state_dict(step.opt, ck.get("optimizer_state_dict"), log)
        det = ck.get("grad_explosion_state")
        if det and hasattr(step.opt, "write_detector_state"):
            step.opt.write_detector_state(float(det["ema_norm"]), int(det["ema_steps"]))
    log(f"resumed from {path} (epoch {int(ck.get('epoch', -1)) + 1})")
    return int(ck.get("epoch", -1)) + 1


def save_phoneme_processor(dataset, output_dir: str, log: Callable[[str], None] = print) -> Optional[str]:
    """<output_dir>/phoneme_processor.pkl = pickle of processor.to_dict(), as the reference trainer writes it at the start of
    training (trainer.py:2828, checkpoint_manager.py:244-249): the reference's resume / inference read the phoneme table from
    that file (checkpoint_manager.py:252-259, 527-530).  The processor itself belongs to the phoneme front-end (out of scope):
    it is whatever object the corpus dataset carries; datasets without one (--synthetic, cached corpora) write nothing."""
    ds = dataset
    while ds is not None and not hasattr(ds, "phoneme_processor"):
        ds = getattr(ds, "ds", None)                # _Subset wrappers of the train / validation split
    proc = getattr(ds, "phoneme_processor", None)
    if proc is None or not hasattr(proc, "to_dict"):
        return None
    import pickle
    path = os.path.join(output_dir, "phoneme_processor.pkl")
    with open(path, "wb") as f:
        pickle.dump(proc.to_dict(), f)
    log(f"phoneme processor saved: {path}")
    return path


def train(cfg: RunConfig, train_ds, val_ds, step, rank: int = 0, world: int = 1,
          log: Callable[[str], None] = print, start_epoch: int = 0) -> Dict:
    """Runs the epochs on an already constructed step object (TrainStep API: micro_step, eval_losses, engine,
    state_dict, store).  Returns a summary dict (per-epoch losses, best epoch, checkpoints written)."""
    from .data import DistributedBatchSampler, collate_fn
    from .engine import AcousticEngine
    hist: List[Dict] = []
    best, best_epoch, since_best, saved = float("inf"), -1, 0, []
    os.makedirs(cfg.output_dir, exist_ok=True)
    if rank == 0:
        save_phoneme_processor(train_ds, cfg.output_dir, log)
    writer = None
    shard = (rank, world) if (cfg.sharded_checkpoints and world > 1) else None
    if cfg.async_checkpoints and (rank == 0 or shard is not None):
        from .checkpoint import AsyncCheckpointWriter
        writer = AsyncCheckpointWriter()
    micro_batches_seen, skipped_seen = 0, 0
    if hasattr(step, "opt") and hasattr(step.opt, "read_ctrl"):
        skipped_seen = int(step.opt.read_ctrl()["skipped_total"])
    for epoch in range(start_epoch, cfg.num_epochs):
        random.seed(cfg.seed + epoch)                 # every rank builds the identical epoch batch list
        sampler = make_sampler(train_ds, cfg)
        batches = list(iter(sampler))
        if world > 1:
            ds_sampler = DistributedBatchSampler(sampler, rank, world, seed=cfg.seed)
            ds_sampler.set_epoch(epoch)
            batches = list(iter(ds_sampler))
        dev_losses = []
        if getattr(train_ds, "thread_safe", False):   # cached corpus: the next batches are read and collated ahead of the step
            from .feature_cache import BatchPrefetcher
            collated = iter(BatchPrefetcher(train_ds, batches, collate_fn))
        else:
            collated = (collate_fn([train_ds[i] for i in b]) for b in batches)
        for win in accumulation_windows(len(batches), cfg.gradient_accumulation_steps):
            for k, bi in enumerate(win):
                batch = next(collated)                # windows cover 0 .. len(batches) - 1 in order
                if epoch >= cfg.spec_augment_start_epoch:
                    B, T = batch["mel_specs"].shape[:2]
                    step.engine.set_spec_augment(AcousticEngine.draw_spec_spans(B, T, step.engine.D))
                else:
                    step.engine.set_spec_augment(None)
                dev_losses.append(step.micro_step(batch, first=(k == 0), last=(k == len(win) - 1),
                                                  divisor=len(win)).clone())
        micro_batches_seen += len(batches)
        rec: Dict = {"epoch": epoch, "batches": len(batches), "global_step": micro_batches_seen}
        # steps the device-side non-finite guard skipped did not update the weights: the reference does not advance its
        # LR schedule on such a step (trainer.py:2479-2482); the host schedule is rolled back here, once per epoch
        if hasattr(step, "opt") and hasattr(step.opt, "read_ctrl") and hasattr(step.sched, "rewind"):
            skipped = int(step.opt.read_ctrl()["skipped_total"])
            if skipped > skipped_seen:
                step.sched.rewind(skipped - skipped_seen)
                log(f"epoch {epoch + 1}: {skipped - skipped_seen} optimizer step(s) skipped (non-finite gradients)")
                skipped_seen = skipped
        if dev_losses:
            mean = torch.stack(dev_losses).mean(dim=0).cpu().tolist()          # one device read per epoch
            rec.update(train_loss=mean[0], train_mel=mean[1], train_dur=mean[2], train_stop=mean[3])
        if val_ds is not None and len(val_ds) and (epoch + 1) % max(1, cfg.validation_interval) == 0:
            vs = make_sampler(val_ds, cfg, shuffle=False)
            acc = step.new_val_metrics() if cfg.val_metrics and hasattr(step, "new_val_metrics") else None
            extra = {} if acc is None else {"metrics": acc}
            vl = [step.eval_losses(collate_fn([val_ds[i] for i in b]), **extra) for b in iter(vs)]
            if vl:
                vm = torch.stack(vl).mean(dim=0).cpu().tolist()
                rec.update(val_loss=vm[0], val_mel_loss=vm[1], val_dur_loss=vm[2], val_stop_loss=vm[3])
                if acc is not None:
                    rec.update({k: v for k, v in step.read_val_metrics(acc).items() if v is not None})
                if vm[0] < best - cfg.early_stopping_min_delta:
                    best, best_epoch, since_best = vm[0], epoch, 0
                    if rank == 0 or shard is not None:
                        saved.append(save_checkpoint(cfg, step, epoch, rec, best, best_epoch, writer, shard))
                else:
                    since_best += 1
        hist.append(rec)
        log(f"epoch {epoch + 1}/{cfg.num_epochs}: " + ", ".join(f"{k}={v:.4f}" for k, v in rec.items()
                                                                  if isinstance(v, float)))
        if (rank == 0 or shard is not None) and cfg.save_every > 0 and (epoch + 1) % cfg.save_every == 0:
            saved.append(save_checkpoint(cfg, step, epoch, rec, best, best_epoch, writer, shard))
        if world > 1:
            # rank 0 alone may just have written a checkpoint (hundreds of MB through torch.save): the others must not
            # enter the next epoch's collective kernel and spin on it for the duration of one-sided host work
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.barrier(device_ids=[torch.cuda.current_device()] if torch.cuda.is_available() else None)
        if val_ds is not None and cfg.early_stopping_patience > 0 and since_best >= cfg.early_stopping_patience:
            log(f"early stopping after epoch {epoch + 1} (best val_loss {best:.4f} at epoch {best_epoch + 1})")
            break
    if writer is not None:
        writer.wait()                                 # every file is on disk (or the failure is raised) before returning
    final = save_final_model(cfg, step) if rank == 0 else None
    return {"history": hist, "best_val_loss": best, "best_val_epoch": best_epoch, "checkpoints": saved, "final_model": final}


def _steps_completed(step) -> int:
    opt = getattr(step, "opt", None)
    if opt is not None and hasattr(opt, "read_ctrl"):
        return int(opt.read_ctrl()["step"])
    return int(step.sched.current_optimizer_step)


def _scheduler_config(step) -> Dict:
    """The hyper-parameters the schedule was built from (the reference stores OneCycleLR's constructor arguments under
    this key and re-anchors on resume, checkpoint_manager.py:757-793)."""
    sc = getattr(step.sched, "cfg", None)
    out = dict(getattr(sc, "__dict__", {}) or {})
    for k in ("base_lr", "max_lr", "warmup_steps", "onecycle_steps", "div_factor"):
        if hasattr(step.sched, k):
            out[k] = getattr(step.sched, k)
    return out


def _detector_state(step) -> Optional[Dict]:
    """Gradient-explosion detector state (norm EMA + the number of norms it has seen, trainer.py:914-925), so that a
    resumed run does not spend min_ema_steps steps without the EMA threshold."""
    opt = getattr(step, "opt", None)
    if opt is None or not hasattr(opt, "read_ctrl"):
        return None
    c = opt.read_ctrl()
    return {"ema_norm": c["ema_norm"], "ema_steps": c["ema_steps"], "skipped_total": c["skipped_total"]}


def save_final_model(cfg: RunConfig, step) -> Optional[str]:
    """<output_dir>/kokoro_russian_final.pth = {model_state_dict, config, model_metadata}, what the reference trainer writes
    when training ends (trainer.py:3013, checkpoint_manager.py:916-926) and what its inference loads first
    (inference/inference.py:111-128)."""
    if not hasattr(step, "state_dict"):
        return None
    path = os.path.join(cfg.output_dir, "kokoro_russian_final.pth")
    payload = {"model_state_dict": {k: v.detach().cpu() for k, v in step.state_dict().items()},
               "config": _picklable_config(cfg)}
    model_cfg = getattr(getattr(step, "engine", None), "cfg", None)
    if model_cfg is not None and hasattr(model_cfg, "hidden_dim"):
        from .checkpoint import build_model_metadata
        payload["model_metadata"] = build_model_metadata(model_cfg, cfg)
    torch.save(payload, path)
    return path


def save_checkpoint(cfg: RunConfig, step, epoch: int, rec: Dict, best: float, best_epoch: int, writer=None,
                    shard: Optional[tuple] = None) -> str:
    """checkpoint_epoch_{N}.pth with the reference's key names (trainer.py:1994-2031) for what this path owns.
    shard = (rank, world): called by every rank, each writes its share (checkpoint.save_sharded)."""
    path = os.path.join(cfg.output_dir, f"checkpoint_epoch_{epoch + 1}.pth")
    st = step.store
    host = (lambda v: v.detach()) if shard is not None else (lambda v: v.detach().cpu())   # sharded: only own tensors are copied
    ckpt = {"epoch": epoch, "model_state_dict": {k: host(v) for k, v in step.state_dict().items()},
            "ema_model_state_dict": ({k: host(v) for k, v in st.state_dict(st.ema).items()}
                                     if getattr(st, "ema", None) is not None else None),
            # counters: scheduler calls made / optimizer steps that really updated the weights (device-side counter: a
            # non-finite step is skipped on the device without a host sync) / micro-batches seen (trainer.py:1994-2031)
            "current_optimizer_step": step.sched.current_optimizer_step,
            "optimizer_steps_completed": _steps_completed(step),
            "global_step": rec.get("global_step", step.sched.current_optimizer_step * max(1, cfg.gradient_accumulation_steps)),
            "ema_updates": _steps_completed(step),
            "scheduler_state_dict": step.sched.state_dict(),
            "scheduler_config": _scheduler_config(step),
            "grad_explosion_state": _detector_state(step),
            "loss": rec.get("train_loss"), "train_loss": rec.get("train_loss"), "val_loss": rec.get("val_loss"),
            "val_mel_loss": rec.get("val_mel_loss"), "val_dur_loss": rec.get("val_dur_loss"),
            "val_stop_loss": rec.get("val_stop_loss"), "best_val_loss": best, "best_val_epoch": best_epoch,
            "config": _picklable_config(cfg)}
    if hasattr(step, "opt"):                          # Adam moments + step in torch.optim.AdamW.state_dict() form
        from .checkpoint import optimizer_state_dict
        ckpt["optimizer_state_dict"] = optimizer_state_dict(step.opt, on_device=shard is not None)
    model_cfg = getattr(getattr(step, "engine", None), "cfg", None)
    if model_cfg is not None and hasattr(model_cfg, "hidden_dim"):
        from .checkpoint import build_model_metadata
        ckpt["model_metadata"] = build_model_metadata(model_cfg, cfg)
    if shard is not None:
        from .checkpoint import save_sharded
        save_sharded(path, ckpt, shard[0], shard[1], writer)
    elif writer is not None:
        writer.save(path, ckpt)                       # serialisation + file system on the writer thread
    else:
        torch.save(ckpt, path)
    return path


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = build_parser().parse_args(argv)
    cfg = create_config_from_args(args)
    noted = [f for f in ("profile_amp", "enable_amp", "disable_amp", "fused_adamw", "no_fused_adamw", "no_mfa")
             if getattr(args, f, False)]
    noted += ["mfa_alignments"] if args.mfa_alignments else []
    if noted:
        print("note: flags without effect on the B200 path: " + ", ".join("--" + f.replace("_", "-") for f in noted),
              file=sys.stderr)
    if not torch.cuda.is_available():
        raise RuntimeError("kokoro-train (B200 path) needs a CUDA device: there is no CPU / MPS fallback")
    from .engine import DropoutConfig
    from .optim import OptimConfig
    from .parallel import init_distributed
    from .params import ModelConfig
    from .train_step import ScheduleConfig, TrainStep
    rank, local, world = init_distributed()
    torch.cuda.set_device(local)
    ds = SyntheticDataset(args.synthetic, seed=cfg.seed) if args.synthetic > 0 else load_reference_dataset(cfg)
    train_ds, val_ds = split_dataset(ds, cfg.validation_split, cfg.seed)
    random.seed(cfg.seed)
    steps_per_epoch = max(1, len(list(iter(make_sampler(train_ds, cfg)))) // max(1, world))
    opt_steps = max(1, (steps_per_epoch + cfg.gradient_accumulation_steps - 1) // cfg.gradient_accumulation_steps)
    # EMA half-life = ema_half_life_epochs (1.0) epochs of optimizer steps, the reference's default (trainer.py:808-822)
    from .optim import recommended_ema_decay
    ema_decay = cfg.ema_decay if cfg.ema_decay is not None else recommended_ema_decay(opt_steps, 1, cfg.ema_half_life_epochs)
    step = TrainStep(ModelConfig(vocab_size=ds.vocab_size), OptimConfig(learning_rate=cfg.learning_rate, ema_decay=ema_decay),
                     ScheduleConfig(total_steps=cfg.num_epochs * opt_steps), device=f"cuda:{local}",
                     process_group=torch.distributed.group.WORLD if world > 1 else None,
                     dropout=DropoutConfig(encoder=cfg.encoder_dropout, decoder=cfg.decoder_dropout,
                                           decoder_input=cfg.decoder_input_dropout, variance=cfg.variance_dropout,
                                           stochastic_depth=cfg.stochastic_depth_rate, seed=cfg.seed + rank))
    step.store.init_default(seed=cfg.seed)
    start_epoch = resume(cfg, step, log=(print if rank == 0 else (lambda s: None)))
    out = train(cfg, train_ds, val_ds, step, rank, world, log=(print if rank == 0 else (lambda s: None)),
                start_epoch=start_epoch)
    if rank == 0:
        print(f"done: {len(out['history'])} epochs, best val_loss {out['best_val_loss']:.4f}, "
              f"{len(out['checkpoints'])} checkpoints in {cfg.output_dir}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
