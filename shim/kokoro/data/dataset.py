"""Collate contract and batch samplers (reference src/kokoro/data/dataset.py:871-1176)."""
from kokoro_ruslan_b200.data import (DistributedBatchSampler, DynamicFrameBatchSampler,  # noqa: F401
                                     LengthBasedBatchSampler, collate_fn)
