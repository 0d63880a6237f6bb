"""tests/golden/sampler.json: batches produced by the LIVE reference DynamicFrameBatchSampler /
LengthBasedBatchSampler / collate_fn for seeded inputs (build container only)."""
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/src")


class DummyDataset:
    def __init__(self, lengths):
        self.samples = [{"audio_length": int(v)} for v in lengths]

    def __len__(self):
        return len(self.samples)


def lengths_for(n, seed):
    rng = np.random.RandomState(seed)
    return np.clip(np.exp(rng.normal(np.log(350.0), 0.55, size=n)), 40, 1800).astype(int).tolist()


def main():
    from kokoro.data.dataset import DynamicFrameBatchSampler, LengthBasedBatchSampler
    out = {"cases": []}
    for n, seed, kw in [(200, 1, dict(max_frames=8000, min_batch_size=4, max_batch_size=32, drop_last=False, shuffle=True)),
                        (57, 2, dict(max_frames=3000, min_batch_size=2, max_batch_size=8, drop_last=True, shuffle=True)),
                        (120, 3, dict(max_frames=20000, min_batch_size=4, max_batch_size=16, drop_last=False, shuffle=False)),
                        (1, 4, dict(max_frames=8000, min_batch_size=1, max_batch_size=4, drop_last=False, shuffle=True))]:
        lens = lengths_for(n, seed)
        random.seed(100 + seed)
        s = DynamicFrameBatchSampler(DummyDataset(lens), **kw)
        first = [list(map(int, b)) for b in s.batches]
        epoch2 = [list(map(int, b)) for b in iter(s)]
        out["cases"].append({"kind": "dynamic", "lengths": lens, "seed": 100 + seed, "kwargs": kw, "init": first,
                             "epoch": epoch2})
    lens = lengths_for(64, 9)
    random.seed(77)
    s = LengthBasedBatchSampler(DummyDataset(lens), batch_size=8, drop_last=False, shuffle=True)
    out["cases"].append({"kind": "length", "lengths": lens, "seed": 77, "kwargs": dict(batch_size=8, drop_last=False, shuffle=True),
                         "epoch": [list(map(int, b)) for b in iter(s)]})
    json.dump(out, open(os.path.join(HERE, "sampler.json"), "w"))
    print("cases", len(out["cases"]), "batches in case 0:", len(out["cases"][0]["epoch"]))


if __name__ == "__main__":
    main()
