"""Per-utterance feature cache in the reference's on-disk format, with read-ahead.

The reference's `RuslanDataset` (src/kokoro/data/dataset.py) keeps one `torch.save`d dict per utterance under
`<corpus>/.feature_cache/<audio_file>.pt` (`_get_feature_cache_path` :412-414, `_save_cached_features` :566-578, payload
:849-862, `FEATURE_CACHE_VERSION = 7` :29) plus a bounded LRU copy in RAM (:416-450), and serves `__getitem__` from it
unless the sample is speed-perturbed (:629-638).  With a ~6 ms optimizer step the per-item `torch.load` on the training
thread is what the loop waits for, so this module

  * reads and writes exactly that format (`FeatureCache.load / save`: same path rule, same keys, same version check, the
    reference's LRU limits by entry count and by estimated bytes) — caches written by either side serve the other;
  * turns the device feature pipeline's batch output (features.FeaturePipeline, N1) into those per-utterance payloads
    (`payloads_from_batch`), so a cold cache is filled at GPU speed;
  * loads the files of the NEXT batches on background threads while the current step runs (`CacheReadAhead`), in the batch
    order the sampler fixed for the epoch;
  * offers the cached corpus as a dataset (`CachedFeatureDataset`) that DynamicFrameBatchSampler / collate_fn consume
    like the reference dataset (`samples[i]["audio_length"]`, item dicts with the reference's keys).

Host-side code only: no kernels, nothing here is on the measured path.
"""
from __future__ import annotations

import threading
import time
from collections import OrderedDict
from concurrent.futures import Future, ThreadPoolExecutor
from pathlib import Path
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import torch

FEATURE_CACHE_VERSION = 7                     # data/dataset.py:29
TENSOR_KEYS = ("mel_spec", "phoneme_indices", "stress_indices", "phoneme_durations", "stop_token_targets", "pitch", "energy")
PAYLOAD_KEYS = TENSOR_KEYS + ("text", "audio_file", "mel_length", "phoneme_length", "_cache_version")   # :849-862


def estimate_feature_size_bytes(features: Dict) -> int:
    """dataset.py:416-424: tensors by storage size, strings by their UTF-8 length, everything else free."""
    total = 0
    for value in features.values():
        if isinstance(value, torch.Tensor):
            total += value.numel() * value.element_size()
        elif isinstance(value, str):
            total += len(value.encode("utf-8"))
    return total


class FeatureCache:
    """Disk cache + bounded RAM LRU, the reference's rules (dataset.py:412-450, 522-578).  Thread-safe: the read-ahead
    workers and the training thread share one instance."""

    def __init__(self, cache_dir, max_entries: int = 30000, max_mb: float = 8192.0, use_memory_cache: bool = True):
        self.cache_dir = Path(cache_dir)
        self.cache_dir.mkdir(parents=True, exist_ok=True)
        self.max_entries = int(max_entries)
        self.max_bytes = int(max(0.0, float(max_mb)) * 1024 * 1024)
        self.use_memory_cache = bool(use_memory_cache)
        self._mem: "OrderedDict[str, Dict]" = OrderedDict()
        self._mem_bytes = 0
        self._lock = threading.Lock()
        self.requests = self.mem_hits = self.disk_hits = self.misses = 0
        self.disk_latency_ns = self.disk_latency_count = 0

    # ---- paths / RAM LRU ---------------------------------------------------------------------------------------------
    def path(self, audio_file: str) -> Path:
        return self.cache_dir / f"{audio_file}.pt"

    def _evict(self) -> None:
        while self._mem:
            over_entries = self.max_entries > 0 and len(self._mem) > self.max_entries
            over_bytes = self.max_bytes > 0 and self._mem_bytes > self.max_bytes
            if not over_entries and not over_bytes:
                break
            _, evicted = self._mem.popitem(last=False)
            self._mem_bytes -= evicted["_cache_mem_bytes"]

    def _put_mem(self, audio_file: str, features: Dict) -> None:
        if audio_file in self._mem:
            self._mem_bytes -= self._mem.pop(audio_file)["_cache_mem_bytes"]
        size = estimate_feature_size_bytes(features)
        self._mem[audio_file] = {"features": features, "_cache_mem_bytes": size}
        self._mem_bytes += size
        self._evict()

    @property
    def memory_entries(self) -> int:
        return len(self._mem)

    @property
    def memory_bytes(self) -> int:
        return self._mem_bytes

    # ---- load / save --------------------------------------------------------------------------------------------------
    def load(self, audio_file: str) -> Optional[Dict]:
        """The cached payload, or None (absent, unreadable, or written by another cache version — dataset.py:522-564)."""
        with self._lock:
            self.requests += 1
            if self.use_memory_cache and audio_file in self._mem:
                entry = self._mem.pop(audio_file)
                if entry["features"].get("_cache_version") == FEATURE_CACHE_VERSION:
                    self._mem[audio_file] = entry                     # most recently used
                    self.mem_hits += 1
                    return entry["features"]
                self._mem_bytes -= entry["_cache_mem_bytes"]          # stale entry: dropped
        p = self.path(audio_file)
        if p.exists():
            try:
                t0 = time.monotonic_ns()
                features = torch.load(p, weights_only=False)
                dt = time.monotonic_ns() - t0
            except Exception:
                features = None
                dt = 0
            if isinstance(features, dict) and features.get("_cache_version") == FEATURE_CACHE_VERSION:
                with self._lock:
                    self.disk_latency_ns += dt
                    self.disk_latency_count += 1
                    self.disk_hits += 1
                    if self.use_memory_cache:
                        self._put_mem(audio_file, features)
                return features
        with self._lock:
            self.misses += 1
        return None

    def save(self, audio_file: str, features: Dict) -> None:
        """Writes the payload (dataset.py:566-578).  `_cache_version` is stamped if the caller left it out."""
        features = dict(features)
        features.setdefault("_cache_version", FEATURE_CACHE_VERSION)
        features.setdefault("audio_file", audio_file)
        p = self.path(audio_file)
        p.parent.mkdir(parents=True, exist_ok=True)
        tmp = p.with_suffix(p.suffix + ".tmp")
        torch.save(features, tmp)
        tmp.replace(p)                                               # readers never see a partial file
        if self.use_memory_cache:
            with self._lock:
                self._put_mem(audio_file, features)

    def stats(self) -> Dict[str, float]:
        with self._lock:
            n = max(1, self.requests)
            return {"requests": self.requests, "mem_hits": self.mem_hits, "disk_hits": self.disk_hits, "misses": self.misses,
                    "hit_rate": (self.mem_hits + self.disk_hits) / n, "memory_entries": len(self._mem),
                    "memory_mb": self._mem_bytes / 2 ** 20,
                    "disk_ms_avg": self.disk_latency_ns / max(1, self.disk_latency_count) / 1e6}


def payloads_from_batch(audio_files: Sequence[str], texts: Sequence[str], mel: torch.Tensor, frames: torch.Tensor,
                        pitch: torch.Tensor, energy: torch.Tensor, phoneme_indices: Sequence[torch.Tensor],
                        stress_indices: Sequence[torch.Tensor], phoneme_durations: Sequence[torch.Tensor],
                        stop_token_targets: Sequence[torch.Tensor]) -> List[Dict]:
    """Per-utterance payloads (dataset.py:849-862) from one batch of the device feature pipeline: mel (B, n_mels, T_max)
    log-mel, frames (B,), pitch / energy (B, T_max) as features.FeaturePipeline returns them; the text-side tensors come from
    the corpus front-end (phonemes, stress, MFA durations, stop targets).  One device-to-host copy per tensor kind."""
    B = len(audio_files)
    mel_h, pitch_h, energy_h = mel.detach().float().cpu(), pitch.detach().float().cpu(), energy.detach().float().cpu()
    n = [int(v) for v in frames.detach().cpu().tolist()]
    out = []
    for b in range(B):
        out.append({
            "mel_spec": mel_h[b, :, :n[b]].clone(),                    # (n_mels, frames), as the reference stores it
            "phoneme_indices": phoneme_indices[b].detach().cpu().long(),
            "stress_indices": stress_indices[b].detach().cpu().long(),
            "phoneme_durations": phoneme_durations[b].detach().cpu().long(),
            "stop_token_targets": stop_token_targets[b].detach().cpu().float(),
            "pitch": pitch_h[b, :n[b]].clone(),
            "energy": energy_h[b, :n[b]].clone(),
            "text": texts[b],
            "audio_file": audio_files[b],
            "mel_length": n[b],
            "phoneme_length": int(phoneme_indices[b].shape[0]),
            "_cache_version": FEATURE_CACHE_VERSION,
        })
    return out


class CacheReadAhead:
    """Loads the cache files of upcoming batches on worker threads while the current optimizer step runs.

    batches: the epoch's batches as lists of audio_file names, in the order the sampler fixed.  Iterating yields, per
    batch, the list of payloads (None for a miss — the caller computes that item and `save`s it).  At most `depth` batches
    are in flight; the RAM LRU of the cache is bypassed for read-ahead loads only in the sense that they populate it."""

    def __init__(self, cache: FeatureCache, batches: Iterable[Sequence[str]], depth: int = 4, workers: int = 4):
        self.cache = cache
        self._batches = iter(batches)
        self.depth = max(1, int(depth))
        self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="kr-cache")
        self._queue: List[List[Future]] = []
        self._closed = False

    def _submit_one(self) -> bool:
        try:
            names = next(self._batches)
        except StopIteration:
            return False
        self._queue.append([self._pool.submit(self.cache.load, n) for n in names])
        return True

    def __iter__(self) -> Iterator[List[Optional[Dict]]]:
        while len(self._queue) < self.depth and self._submit_one():
            pass
        while self._queue:
            futures = self._queue.pop(0)
            self._submit_one()                                        # keep `depth` batches in flight
            yield [f.result() for f in futures]
        self.close()

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._pool.shutdown(wait=False, cancel_futures=True)


class CachedFeatureDataset(torch.utils.data.Dataset):
    """The cached part of a corpus as a map-style dataset with the reference dataset's surface for the samplers and the
    collate function: `samples[i]` = {"audio_file", "text", "audio_length" (mel frames)}, `dataset[i]` = the payload dict
    with `text` / `audio_file` refreshed (dataset.py:632-638)."""

    def __init__(self, cache: FeatureCache, samples: Sequence[Dict]):
        self.cache = cache
        self.samples = [dict(s) for s in samples]

    @classmethod
    def scan(cls, cache: FeatureCache, texts: Optional[Dict[str, str]] = None) -> "CachedFeatureDataset":
        """Every current-version file in the cache directory (sorted by name); lengths come from the payloads."""
        samples = []
        for p in sorted(cache.cache_dir.glob("*.pt")):
            name = p.name[:-3]
            f = cache.load(name)
            if f is not None:
                samples.append({"audio_file": name, "text": (texts or {}).get(name, f.get("text", "")),
                                "audio_length": int(f["mel_length"])})
        return cls(cache, samples)

    def __len__(self) -> int:
        return len(self.samples)

    def __getitem__(self, idx: int) -> Dict:
        s = self.samples[idx]
        f = self.cache.load(s["audio_file"])
        if f is None:
            raise KeyError(f"{s['audio_file']}: not in the feature cache {self.cache.cache_dir} (version {FEATURE_CACHE_VERSION})")
        item = dict(f)                                               # never hand out the shared RAM copy itself
        item["text"] = s["text"]
        item["audio_file"] = s["audio_file"]
        return item


class BatchPrefetcher:
    """Collated batches `depth` ahead of the training loop: the items of upcoming batches are fetched (cache reads) and
    collated on worker threads while the current optimizer step runs.  Only for datasets that declare `thread_safe`
    (CachedFeatureDataset and views over it); yields exactly `collate([dataset[i] for i in batch])` in the given order."""

    def __init__(self, dataset, batches: Sequence[Sequence[int]], collate, depth: int = 3, workers: int = 4):
        if not getattr(dataset, "thread_safe", False):
            raise ValueError("BatchPrefetcher needs a dataset that declares thread_safe = True")
        self.dataset, self.batches, self.collate = dataset, [list(b) for b in batches], collate
        self.depth, self.workers = max(1, int(depth)), max(1, int(workers))

    def __len__(self) -> int:
        return len(self.batches)

    def __iter__(self):
        def build(idxs):
            return self.collate([self.dataset[i] for i in idxs])
        pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="kr-batch")
        try:
            pending: List[Future] = []
            nxt = 0
            while nxt < len(self.batches) and len(pending) < self.depth:
                pending.append(pool.submit(build, self.batches[nxt]))
                nxt += 1
            while pending:
                fut = pending.pop(0)
                if nxt < len(self.batches):
                    pending.append(pool.submit(build, self.batches[nxt]))
                    nxt += 1
                yield fut.result()
        finally:
            pool.shutdown(wait=False, cancel_futures=True)


CachedFeatureDataset.thread_safe = True
