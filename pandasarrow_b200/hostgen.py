"""Host (numpy) twin of the device-side synthetic generator (csrc/emit.cuh k_synth_*), SURVEY.md §8d.
Data generation only — used for host-buffer (e2e) runs, the CPU baseline and the tests."""
from __future__ import annotations

import numpy as np

SEED_K, SEED_V, SEED_N, SEED_T = 42, 1337, 7, 99
T0_NS = 1577836800 * 10**9
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _rows(n, first_row):
    return np.arange(first_row, first_row + n, dtype=np.uint64)


def keys(n: int, n_groups: int, first_row: int = 0, seed: int = SEED_K) -> np.ndarray:
    return (splitmix64(_rows(n, first_row) ^ np.uint64(seed)) % np.uint64(n_groups)).astype(np.int64)


def vals(n: int, first_row: int = 0, seed: int = SEED_V) -> np.ndarray:
    with np.errstate(over="ignore"):
        r = splitmix64(_rows(n, first_row) + np.uint64(seed))
    return (r >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def valid_mask(n: int, first_row: int = 0, seed: int = SEED_N, null_every: int = 10) -> np.ndarray:
    with np.errstate(over="ignore"):
        r = splitmix64(_rows(n, first_row) + np.uint64(seed))
    return (r % np.uint64(null_every)) != 0


def timestamps(n: int, first_row: int = 0, t0_ns: int = T0_NS, step_ns: int = 60_000_000, seed: int = SEED_T) -> np.ndarray:
    rows = _rows(n, first_row)
    with np.errstate(over="ignore"):
        jitter = splitmix64(rows + np.uint64(seed)) % np.uint64(step_ns)
    return (np.int64(t0_ns) + rows.astype(np.int64) * np.int64(step_ns) + jitter.astype(np.int64))
