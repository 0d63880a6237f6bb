"""Batch assembly for the training hot path — host-side mirror of the reference's
``collate_fn`` (src/kokoro/data/dataset.py:871-921), ``DynamicFrameBatchSampler`` (:924-1142) and
``LengthBasedBatchSampler`` (:1145-1176), plus the rank-aware wrapper the reference does not have
(it is single-process).

Same contracts: ``collate_fn`` returns the 9 tensors + 2 lists the trainer expects, zero padded to
the batch maxima; the samplers read only ``dataset.samples[i]['audio_length']`` / ``len(dataset)``,
expose ``.batches`` and rebuild on every ``__iter__`` when shuffling, drawing from the global
``random`` module in the same order as the reference (so a given ``random.seed`` produces the same
epoch).  B200-first difference: collated tensors are allocated in PINNED host memory so that
``TrainStep.stage`` can issue asynchronous H2D copies.
"""
from __future__ import annotations

import random
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch

_TENSOR_KEYS = ("mel_specs", "phoneme_indices", "stress_indices", "phoneme_durations", "stop_token_targets",
                "pitches", "energies", "mel_lengths", "phoneme_lengths")


def _host_zeros(shape, dtype, pinned: bool) -> torch.Tensor:
    t = torch.zeros(shape, dtype=dtype)
    if pinned:
        try:
            t = t.pin_memory()
        except RuntimeError:      # no CUDA runtime in this process
            pass
    return t


def collate_fn(batch: List[Dict], pin_memory: Optional[bool] = None) -> Dict:
    """list of dataset items -> padded batch dict (reference dataset.py:871-921).

    Item keys used: ``mel_spec (n_mels, T)``, ``pitch (T,)``, ``energy (T,)``, ``stop_token_targets (T,)``,
    ``phoneme_indices / phoneme_durations / stress_indices (P,)``, ``mel_length``, ``phoneme_length``,
    ``text``, ``audio_file``."""
    if pin_memory is None:
        pin_memory = torch.cuda.is_available()
    n = len(batch)
    t_len = [int(it["mel_length"]) for it in batch]
    p_len = [int(it["phoneme_length"]) for it in batch]
    T, P = max(t_len), max(p_len)
    n_mels = batch[0]["mel_spec"].shape[0]
    out = {
        "mel_specs": _host_zeros((n, T, n_mels), torch.float32, pin_memory),
        "phoneme_indices": _host_zeros((n, P), torch.long, pin_memory),
        "stress_indices": _host_zeros((n, P), torch.long, pin_memory),
        "phoneme_durations": _host_zeros((n, P), torch.long, pin_memory),
        "stop_token_targets": _host_zeros((n, T), torch.float32, pin_memory),
        "pitches": _host_zeros((n, T), torch.float32, pin_memory),
        "energies": _host_zeros((n, T), torch.float32, pin_memory),
    }
    frame_fields = (("pitches", "pitch"), ("energies", "energy"), ("stop_token_targets", "stop_token_targets"))
    token_fields = (("phoneme_indices", "phoneme_indices"), ("phoneme_durations", "phoneme_durations"),
                    ("stress_indices", "stress_indices"))
    for row, (item, tl, pl) in enumerate(zip(batch, t_len, p_len)):
        out["mel_specs"][row, :tl].copy_(item["mel_spec"].transpose(0, 1)[:tl])
        for dst, src in frame_fields:
            out[dst][row, :tl].copy_(item[src][:tl])
        for dst, src in token_fields:
            out[dst][row, :pl].copy_(item[src][:pl])
    out["mel_lengths"] = torch.tensor(t_len, dtype=torch.long)
    out["phoneme_lengths"] = torch.tensor(p_len, dtype=torch.long)
    out["texts"] = [it["text"] for it in batch]
    out["audio_files"] = [it["audio_file"] for it in batch]
    return out


class DynamicFrameBatchSampler(torch.utils.data.Sampler):
    """Frame-budget batching (reference dataset.py:924-1142).

    1. samples are split into <= 16 (sqrt(N)) quantile buckets of similar length;
    2. inside a bucket (shuffled) samples are packed greedily while
       ``(n + 1) * max_len <= max_frames`` and ``n < max_batch_size``; a batch shorter than
       ``min_batch_size`` is dropped only if ``drop_last``;
    3. with shuffling, the ``max(2, floor(sqrt(n_batches)))`` costliest batches become evenly spaced
       anchors (heaviest first) and the shuffled remaining batches fill the gaps between them.
    """

    def __init__(self, dataset, max_frames: int = 20000, min_batch_size: int = 4, max_batch_size: int = 32,
                 drop_last: bool = False, shuffle: bool = True):
        self.dataset = dataset
        self.max_frames = max_frames
        self.min_batch_size = min_batch_size
        self.max_batch_size = max_batch_size
        self.drop_last = drop_last
        self.shuffle = shuffle
        self.batches = self._create_batches()

    def _get_sample_frames(self, idx: int) -> int:
        return self.dataset.samples[idx]["audio_length"]

    def _keep(self, group: List[int]) -> bool:
        return bool(group) and (len(group) >= self.min_batch_size or not self.drop_last)

    def _pack(self, members: Sequence[int], frames: np.ndarray) -> List[List[int]]:
        packed: List[List[int]] = []
        cur: List[int] = []
        longest = 0
        for idx in members:
            f = int(frames[idx])
            if cur and ((len(cur) + 1) * max(longest, f) > self.max_frames or len(cur) >= self.max_batch_size):
                if self._keep(cur):
                    packed.append(cur)
                cur, longest = [], 0
            cur.append(idx)
            longest = max(longest, f)
        if self._keep(cur):
            packed.append(cur)
        return packed

    def _create_batches(self) -> List[List[int]]:
        n = len(self.dataset)
        if n == 0:
            return []
        frames = np.fromiter((self._get_sample_frames(i) for i in range(n)), dtype=np.int64, count=n)
        n_buckets = min(16, max(1, int(np.sqrt(n))))
        edges = np.percentile(frames, np.linspace(0, 100, n_buckets + 1))
        which = np.clip(np.searchsorted(edges, frames, side="right") - 1, 0, n_buckets - 1)
        batches: List[List[int]] = []
        for b in range(n_buckets):
            members = np.nonzero(which == b)[0].tolist()
            if not members:
                continue
            if self.shuffle:
                random.shuffle(members)
            batches.extend(self._pack(members, frames))
        if not self.shuffle or len(batches) <= 1:
            return batches
        # heavy-batch spreading
        cost = [int(frames[b].max()) * len(b) for b in batches]
        ranked = [batches[i] for i in sorted(range(len(batches)), key=cost.__getitem__, reverse=True)]
        n_anchor = max(2, int(len(batches) ** 0.5))
        anchors, rest = ranked[:n_anchor], ranked[n_anchor:]
        random.shuffle(rest)
        base, extra = divmod(len(rest), n_anchor)
        spread: List[List[int]] = []
        pos = 0
        for k, anchor in enumerate(anchors):
            take = base + (1 if k < extra else 0)
            spread.append(anchor)
            spread.extend(rest[pos:pos + take])
            pos += take
        return spread

    def __iter__(self) -> Iterator[List[int]]:
        if self.shuffle:
            self.batches = self._create_batches()
        yield from self.batches

    def __len__(self) -> int:
        return len(self.batches)


class LengthBasedBatchSampler(torch.utils.data.Sampler):
    """Fixed batch size, length-grouped (reference dataset.py:1145-1176): a DynamicFrameBatchSampler whose
    frame budget can never bind."""

    def __init__(self, dataset, batch_size: int, drop_last: bool = False, shuffle: bool = True):
        self.dataset, self.batch_size, self.drop_last, self.shuffle = dataset, batch_size, drop_last, shuffle
        longest = max((dataset.samples[i]["audio_length"] for i in range(len(dataset))), default=10000)
        self._delegate = DynamicFrameBatchSampler(dataset, max_frames=longest * batch_size, min_batch_size=1,
                                                  max_batch_size=batch_size, drop_last=drop_last, shuffle=shuffle)

    @property
    def batches(self):
        return self._delegate.batches

    def __iter__(self):
        yield from self._delegate

    def __len__(self) -> int:
        return len(self._delegate)


class DistributedBatchSampler(torch.utils.data.Sampler):
    """Data-parallel view of a batch sampler (new: the reference has no multi-process path).

    Every rank re-seeds the global ``random`` module with ``seed + epoch`` and builds the IDENTICAL epoch
    batch list, then takes batches ``rank, rank + world, ...`` truncated to the same count on every rank,
    so all ranks run the same number of optimizer steps and the per-step gradient all-reduce pairs up.
    """

    def __init__(self, sampler, rank: int, world_size: int, seed: int = 0):
        if not 0 <= rank < world_size:
            raise ValueError("rank must be in [0, world_size)")
        self.sampler, self.rank, self.world_size, self.seed = sampler, rank, world_size, seed
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _epoch_batches(self) -> List[List[int]]:
        state = random.getstate()
        random.seed(self.seed + self.epoch)
        try:
            everything = list(iter(self.sampler))
        finally:
            random.setstate(state)
        usable = len(everything) // self.world_size * self.world_size
        return everything[self.rank:usable:self.world_size]

    def __iter__(self):
        yield from self._epoch_batches()

    def __len__(self) -> int:
        return len(self.sampler) // self.world_size


def build_stop_token_targets(length: int, tail: int = 4, decay: float = 0.5) -> torch.Tensor:
    """Soft stop ramp frame[T-1-k] = decay^k for k <= tail (reference dataset.py:32-64)."""
    t = torch.zeros(max(int(length), 0))
    for k in range(min(int(tail) + 1, t.numel())):
        t[t.numel() - 1 - k] = decay ** k
    return t
