#!/usr/bin/env python
"""One batched RecoverPolyFromSamples call (config 4 shape: n = 2^14, half missing, 64 polynomials) and one batched
DASFFTExtension, for a per-kernel launch list."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go_kzg_b200 as kzg                       # noqa: E402
from go_kzg_b200.synth import random_fr_limbs   # noqa: E402

scale, batch = 14, 64
n = 1 << scale
fs = kzg.FFTSettings(scale)
even = np.stack([random_fr_limbs(n // 2, 100 + b) for b in range(batch)])
odd = fs.das_fft_extension_batch(even)
full = np.empty((batch, n, 4), dtype=np.uint64)
full[:, 0::2], full[:, 1::2] = even, odd
rng = np.random.default_rng(14)
present = np.ones((batch, n), dtype=np.uint8)
for b in range(batch):
    present[b, rng.permutation(n)[: n // 2]] = 0
samples = full.copy()
samples[present == 0] = 0
rec = fs.recover_poly_from_samples_batch(samples, present)
assert np.array_equal(rec, full)
rec = fs.recover_poly_from_samples_batch(samples, present)
