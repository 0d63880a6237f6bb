"""Writes tests/golden/feature_cache/utt_ref.pt with the LIVE reference's own cache writer
(RuslanDataset._save_cached_features, src/kokoro/data/dataset.py:566-578) on a synthetic payload with the reference's keys
(:849-862), and checks that the reference's loader reads it back.  Run in the build container:
    python tests/golden/make_golden_feature_cache.py"""
import logging
import os
import sys
from collections import OrderedDict
from pathlib import Path

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_trainer import _import_reference  # noqa: E402

_import_reference()
logging.disable(logging.WARNING)
from kokoro.data import dataset as rd  # noqa: E402


def stub_dataset(cache_dir: Path):
    ds = rd.RuslanDataset.__new__(rd.RuslanDataset)
    ds.use_feature_cache = True
    ds.use_memory_cache = True
    ds.feature_cache_dir = cache_dir
    ds.feature_cache = OrderedDict()
    ds.feature_cache_total_bytes = 0
    ds.feature_cache_max_entries = 4
    ds.feature_cache_max_bytes = 1 << 30
    ds.feature_cache_mem_latency_ns = ds.feature_cache_mem_latency_count = 0
    ds.feature_cache_disk_latency_ns = ds.feature_cache_disk_latency_count = 0
    ds.feature_cache_mem_hits = ds.feature_cache_disk_hits = ds.feature_cache_misses = 0
    return ds


def payload(seed: int, frames: int = 37, phonemes: int = 9):
    g = torch.Generator().manual_seed(seed)
    dur = torch.ones(phonemes, dtype=torch.long) * (frames // phonemes)
    dur[-1] += frames - int(dur.sum())
    stop = torch.zeros(frames)
    stop[-4:] = torch.tensor([0.125, 0.25, 0.5, 1.0])
    return {"mel_spec": torch.randn(80, frames, generator=g), "phoneme_indices": torch.randint(1, 59, (phonemes,), generator=g),
            "stress_indices": torch.randint(0, 3, (phonemes,), generator=g), "phoneme_durations": dur, "stop_token_targets": stop,
            "pitch": torch.rand(frames, generator=g), "energy": torch.rand(frames, generator=g), "text": "привет, мир",
            "audio_file": "utt_ref", "mel_length": frames, "phoneme_length": phonemes, "_cache_version": rd.FEATURE_CACHE_VERSION}


if __name__ == "__main__":
    out = Path(HERE) / "feature_cache"
    out.mkdir(exist_ok=True)
    ds = stub_dataset(out)
    ds._save_cached_features("utt_ref", payload(0))
    ds.feature_cache.clear()
    back = ds._load_cached_features("utt_ref")
    assert back is not None and torch.equal(back["mel_spec"], payload(0)["mel_spec"])
    assert str(ds._get_feature_cache_path("a/b")) == str(out / "a/b.pt")
    print("wrote", out / "utt_ref.pt", "version", rd.FEATURE_CACHE_VERSION, "keys", sorted(back))
