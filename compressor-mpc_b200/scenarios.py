"""Deterministic synthetic plant scenarios, identical for the GPU path and its CPU checker.

Scenario 0 is the nominal one of the setup file (so the reference's recorded trajectory
applies).  For s > 0 a counter-based generator (splitmix64 of seed ^ s) perturbs the
initial state by 1 %, the disturbance amplitude by 50 % and its onset by +-200 records.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EEDC0DE
MASK = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(MASK)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(MASK)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(MASK)
    return z ^ (z >> np.uint64(31))


def _uniform(s: np.ndarray, stream: int) -> np.ndarray:
    """U(-1, 1) for scenario indices s and a stream id."""
    with np.errstate(over="ignore"):
        k = _splitmix64(np.uint64(SEED) ^ (s.astype(np.uint64) * np.uint64(1024) + np.uint64(stream)))
    return (k >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def make_scenarios(setup, x_default: np.ndarray, batch: int, n_steps: int, first: int = 0, Ts: float = 0.05):
    """Returns x0 (B, n), block_end (B, n_blocks) int32, block_off (B, n_blocks, n_inputs).

    `first` is the global index of the first scenario (for sharding across ranks)."""
    s = np.arange(first, first + batch)
    n = x_default.shape[0]
    x0 = np.empty((batch, n))
    for i in range(n):
        x0[:, i] = x_default[i] * (1.0 + 0.01 * _uniform(s, i))
    nb = setup.sim_offsets.shape[0]
    ends = setup.block_end_records(Ts=Ts).astype(np.int64)
    block_end = np.tile(ends, (batch, 1))
    block_off = np.tile(setup.sim_offsets[None], (batch, 1, 1)).astype(np.float64)
    amp = 1.0 + 0.5 * _uniform(s, 100)
    shift = np.rint(200.0 * _uniform(s, 101)).astype(np.int64)
    if nb >= 2:
        block_off[:, 1:, :] *= amp[:, None, None]
        block_end[:, 0] = ends[0] + shift
    nominal = s == 0
    x0[nominal] = x_default
    block_off[nominal] = setup.sim_offsets
    block_end[nominal] = ends
    block_end = np.minimum(block_end, n_steps)
    block_end[:, -1] = n_steps
    return x0, np.ascontiguousarray(block_end, dtype=np.int32), np.ascontiguousarray(block_off)
