"""ORACLE (test infrastructure, never the product path) — numpy restatement of the counter-based
dropout mask generator of kokoro_ruslan_b200/csrc/kr_common.cuh (drop_key / drop_hash / drop_pair).

New functionality without a reference counterpart (the reference draws its masks from torch's Philox
stream, which cannot be reproduced): the specification is the header comment of kr_common.cuh
   key(seed, step, site) = splitmix64 finaliser of seed + GOLD*(step+1) + C*(site+1)  -> (k0, k1) 32-bit words
   hash(pair)            = x = pair ^ k0; x *= 0x7feb352d; x ^= x >> 15; x ^= k1; x *= 0x846ca68b; x ^= x >> 16
   keep(e)               = 16-bit lane (e & 1) of hash(e >> 1) >= round(p * 65536)
tests/test_dropout_gpu.py checks the CUDA masks against this file bit for bit.
"""
import numpy as np

_M64 = (1 << 64) - 1


def drop_key(seed: int, step: int, site: int):
    z = (seed + 0x9E3779B97F4A7C15 * (step + 1) + 0xD1B54A32D192ED03 * (site + 1)) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    z ^= z >> 31
    return np.uint32(z & 0xFFFFFFFF), np.uint32(z >> 32)


def drop_thr(p: float) -> int:
    return max(0, min(65535, int(round(float(p) * 65536.0))))


def drop_thr8(p: float) -> int:
    return max(0, min(255, int(round(float(p) * 256.0)))) * 256


def keep_mask(seed: int, step: int, site: int, p: float, rows: int, cols: int, ld: int = None,
              byte_lanes: bool = False) -> np.ndarray:
    """uint8 [rows, cols]: 1 = kept, for elements r * ld + c.  byte_lanes: the attention-probability sites'
    mode (hash(e >> 2), 8-bit lane e & 3, threshold round(p * 256))."""
    ld = cols if ld is None else ld
    k0, k1 = drop_key(seed, step, site)
    e = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(ld) + np.arange(cols, dtype=np.uint64)[None, :])
    x = (e >> np.uint64(2 if byte_lanes else 1)).astype(np.uint32) ^ k0
    with np.errstate(over="ignore"):
        x = x * np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x ^= k1
        x = x * np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    if byte_lanes:
        lane = (x >> ((e & np.uint64(3)) * np.uint64(8)).astype(np.uint32)) & np.uint32(0xFF)
        return (lane >= np.uint32(drop_thr8(p) >> 8)).astype(np.uint8)
    lane = np.where((e & np.uint64(1)) == 1, x >> np.uint32(16), x & np.uint32(0xFFFF))
    return (lane >= np.uint32(drop_thr(p))).astype(np.uint8)
