"""Feature cache in the reference's on-disk format (reference src/kokoro/data/dataset.py:412-578,
849-866): a file written by the LIVE reference's writer (tests/golden/feature_cache/utt_ref.pt, made by
tests/golden/make_golden_feature_cache.py) is served, files written here carry the same keys / version / path rule (and
are read back by the live reference's loader when baseline/_ref is present), the RAM LRU follows the reference's limits,
the read-ahead yields batches in order, and the cached corpus feeds the samplers and collate_fn."""
import os
import shutil
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

from kokoro_ruslan_b200 import feature_cache as fc  # noqa: E402


def _payload(name, frames, phonemes=7, seed=0):
    g = torch.Generator().manual_seed(seed)
    dur = torch.ones(phonemes, dtype=torch.long) * (frames // phonemes)
    dur[-1] += frames - int(dur.sum())
    return {"mel_spec": torch.randn(80, frames, generator=g), "phoneme_indices": torch.randint(1, 59, (phonemes,), generator=g),
            "stress_indices": torch.randint(0, 3, (phonemes,), generator=g), "phoneme_durations": dur,
            "stop_token_targets": torch.zeros(frames), "pitch": torch.rand(frames, generator=g),
            "energy": torch.rand(frames, generator=g), "text": f"text of {name}", "audio_file": name, "mel_length": frames,
            "phoneme_length": phonemes}


def test_reads_a_file_written_by_the_reference(tmp_path):
    shutil.copy(os.path.join(HERE, "golden", "feature_cache", "utt_ref.pt"), tmp_path / "utt_ref.pt")
    cache = fc.FeatureCache(tmp_path)
    f = cache.load("utt_ref")
    assert f is not None and set(f) == set(fc.PAYLOAD_KEYS) and f["_cache_version"] == fc.FEATURE_CACHE_VERSION == 7
    assert f["mel_spec"].shape == (80, 37) and f["mel_length"] == 37 and f["text"] == "привет, мир"
    assert int(f["phoneme_durations"].sum()) == 37
    assert cache.load("utt_ref") is f                       # second request: the RAM copy
    s = cache.stats()
    assert (s["disk_hits"], s["mem_hits"], s["misses"]) == (1, 1, 0)
    assert cache.load("absent") is None and cache.stats()["misses"] == 1


def test_written_files_have_the_reference_format_and_stale_versions_are_misses(tmp_path):
    cache = fc.FeatureCache(tmp_path, use_memory_cache=False)
    cache.save("sub/utt_a", _payload("sub/utt_a", 50))
    assert (tmp_path / "sub" / "utt_a.pt").exists()          # <dir>/<audio_file>.pt, dataset.py:412-414
    raw = torch.load(tmp_path / "sub" / "utt_a.pt", weights_only=False)
    assert set(raw) == set(fc.PAYLOAD_KEYS) and raw["_cache_version"] == 7
    stale = dict(raw, _cache_version=6)
    torch.save(stale, tmp_path / "old.pt")
    assert cache.load("old") is None                         # dataset.py:553-554
    (tmp_path / "broken.pt").write_bytes(b"not a pickle")
    assert cache.load("broken") is None                      # dataset.py:560-562: unreadable = miss, no exception
    ref_dir = os.path.join(ROOT, "baseline", "_ref", "kokoro")
    if os.path.isdir(ref_dir):                               # the live reference's loader reads what we wrote
        import make_golden_feature_cache as mg
        ds = mg.stub_dataset(tmp_path)
        back = ds._load_cached_features("sub/utt_a")
        assert back is not None and torch.equal(back["mel_spec"], raw["mel_spec"]) and back["mel_length"] == 50


def test_memory_lru_limits_follow_the_reference(tmp_path):
    cache = fc.FeatureCache(tmp_path, max_entries=2, max_mb=1024)
    for i, name in enumerate(("a", "b", "c")):
        cache.save(name, _payload(name, 20 + i))
    assert cache.memory_entries == 2                         # "a" (least recently used) was evicted
    cache.load("b")
    cache.save("d", _payload("d", 30))
    assert cache.memory_entries == 2 and cache.load("b") is not None and cache.stats()["mem_hits"] >= 2
    one = fc.estimate_feature_size_bytes(dict(_payload("x", 100), _cache_version=7))
    assert one == (80 * 100 + 100 * 3) * 4 + 7 * 8 * 3 + len("text of x") + 1     # tensors + utf-8 strings only
    small = fc.FeatureCache(tmp_path, max_entries=0, max_mb=one * 1.5 / 2 ** 20)  # byte cap: one such payload fits, not two
    small.save("p", _payload("p", 100))
    small.save("q", _payload("q", 100))
    assert small.memory_entries == 1


def test_read_ahead_yields_batches_in_order_and_marks_misses(tmp_path):
    cache = fc.FeatureCache(tmp_path, use_memory_cache=False)
    names = [f"u{i:02d}" for i in range(10)]
    for i, n in enumerate(names):
        if n != "u07":
            cache.save(n, _payload(n, 10 + i, seed=i))
    batches = [names[0:3], names[3:6], names[6:10]]
    got = list(fc.CacheReadAhead(cache, batches, depth=2, workers=3))
    assert [len(b) for b in got] == [3, 3, 4]
    for want, items in zip(batches, got):
        for n, f in zip(want, items):
            assert (f is None) if n == "u07" else (f["audio_file"] == n and f["mel_length"] == 10 + names.index(n))


def test_cached_corpus_feeds_sampler_and_collate(tmp_path):
    from kokoro_ruslan_b200.data import DynamicFrameBatchSampler, collate_fn
    cache = fc.FeatureCache(tmp_path)
    frames = [40, 55, 70, 90, 120, 64, 33, 81]
    B = len(frames)
    # payloads_from_batch: one padded device-pipeline batch -> per-utterance payloads (here from CPU tensors)
    T = max(frames)
    mel = torch.randn(B, 80, T)
    pitch, energy = torch.rand(B, T), torch.rand(B, T)
    ph = [torch.randint(1, 59, (6 + b,)) for b in range(B)]
    durs = []
    for b in range(B):
        d = torch.ones(6 + b, dtype=torch.long) * (frames[b] // (6 + b))
        d[-1] += frames[b] - int(d.sum())
        durs.append(d)
    items = fc.payloads_from_batch([f"utt{b}" for b in range(B)], [f"t{b}" for b in range(B)], mel, torch.tensor(frames), pitch,
                                   energy, ph, [torch.zeros_like(p) for p in ph], durs, [torch.zeros(n) for n in frames])
    for it in items:
        assert it["mel_spec"].shape == (80, it["mel_length"]) and it["pitch"].shape == (it["mel_length"],)
        cache.save(it["audio_file"], it)
    ds = fc.CachedFeatureDataset.scan(cache)
    assert len(ds) == B and [s["audio_length"] for s in ds.samples] == frames
    sampler = DynamicFrameBatchSampler(ds, max_frames=300, min_batch_size=1, max_batch_size=4, shuffle=False)
    seen = []
    for idxs in sampler:
        batch = collate_fn([ds[i] for i in idxs], pin_memory=False)
        assert batch["mel_specs"].shape[0] == len(idxs) and batch["mel_specs"].shape[2] == 80
        assert batch["mel_lengths"].tolist() == [frames[i] for i in idxs]
        seen += idxs
    assert sorted(seen) == list(range(B))
    with pytest.raises(KeyError):
        fc.CachedFeatureDataset(cache, [{"audio_file": "nope", "text": "", "audio_length": 1}])[0]
