"""Synthetic benchmark inputs: seeded uniform Fr polynomials.

The reference draws test scalars from crypto/rand (bls/bignum_kilic.go:86-93), which is not
reproducible, so the generator is pinned here: a splitmix64 stream with seed
0xB2000000 + blob index, 4 x u64 little-endian per draw, top bit cleared, rejected if >= r."""
from __future__ import annotations

import numpy as np

_R_LIMBS = np.array([0xffffffff00000001, 0x53bda402fffe5bfe, 0x3339d80809a1d805, 0x73eda753299d7d48], dtype=np.uint64)


def _splitmix64_block(start_state: int, count: int) -> np.ndarray:
    """outputs of splitmix64 for states start + k * GAMMA, k = 1..count"""
    gamma = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(start_state) + gamma * np.arange(1, count + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _below_r(v: np.ndarray) -> np.ndarray:
    lt = np.zeros(v.shape[0], dtype=bool)
    eq = np.ones(v.shape[0], dtype=bool)
    for j in (3, 2, 1, 0):
        lt |= eq & (v[:, j] < _R_LIMBS[j])
        eq &= v[:, j] == _R_LIMBS[j]
    return lt


def random_fr_limbs(n: int, seed: int) -> np.ndarray:
    """(n, 4) uint64 canonical little-endian limbs, uniform in [0, r)."""
    out = np.zeros((0, 4), dtype=np.uint64)
    state = seed & 0xFFFFFFFFFFFFFFFF
    while out.shape[0] < n:
        draws = max(16, int((n - out.shape[0]) * 1.2) + 8)
        w = _splitmix64_block(state, 4 * draws).reshape(draws, 4)
        state = (state + 4 * draws * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        w[:, 3] &= np.uint64((1 << 63) - 1)
        out = np.concatenate([out, w[_below_r(w)]])
    return np.ascontiguousarray(out[:n])


def blob_polys(batch: int, n: int, first_blob: int = 0) -> np.ndarray:
    """(batch, n, 4) uint64: polynomial b uses seed 0xB2000000 + first_blob + b."""
    return np.stack([random_fr_limbs(n, 0xB2000000 + first_blob + b) for b in range(batch)])
