#!/usr/bin/env python
"""One generic LinCombG1 at n = 2^20 (and 2^16) for a per-kernel launch list of the bucket MSM at scale."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go_kzg_b200 as kzg                       # noqa: E402
from go_kzg_b200.synth import random_fr_limbs   # noqa: E402

raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "trusted_setup_g1.bin"), dtype=np.uint8).reshape(2, 4096, 48)
pts = kzg.g1_from_compressed_device(raw[0])
for lg in (16, 20):
    n = 1 << lg
    big = np.concatenate([pts] * (n // 4096))
    sc = random_fr_limbs(n, 7 + lg)
    kzg.lincomb_g1(big, sc)
