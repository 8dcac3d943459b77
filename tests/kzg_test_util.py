"""Shared helpers for the parity tests: the pinned synthetic-input generator of SURVEY.md
section 8d and a few conversions."""
import numpy as np

R = 52435875175126190479447740508185965837690552500527637822603658699938581184513
_M64 = (1 << 64) - 1


def _splitmix64(state):
    state = (state + 0x9E3779B97F4A7C15) & _M64
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return state, z ^ (z >> 31)


def random_fr_ints(n, seed):
    """n uniform field elements: splitmix64 stream, 4 x u64 little-endian, top bit cleared,
    rejection-sampled below r (seed = 0xB2000000 + blob index for the benchmark blobs)."""
    out = []
    st = seed & _M64
    while len(out) < n:
        v = 0
        for j in range(4):
            st, w = _splitmix64(st)
            v |= w << (64 * j)
        v &= (1 << 255) - 1
        if v < R:
            out.append(v)
    return out


def random_fr_limbs(n, seed):
    """Vectorised equivalent of random_fr_ints -> (n, 4) uint64 (same stream, same rejection)."""
    from go_kzg_b200.synth import random_fr_limbs as impl
    return impl(n, seed)


def blob_polys(batch, n, first_blob=0):
    return np.stack([random_fr_limbs(n, 0xB2000000 + first_blob + b) for b in range(batch)])
