#!/usr/bin/env python
"""Two generic LinCombG1 calls (n = 4096 over the trusted setup's points, no table) -- the workload ncu captures
for the bucket-MSM kernels, and a wall-clock figure for the whole host-buffer call."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go_kzg_b200 as kzg                       # noqa: E402
from go_kzg_b200.synth import random_fr_limbs   # noqa: E402

raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "trusted_setup_g1.bin"), dtype=np.uint8).reshape(2, 4096, 48)
pts = kzg.g1_from_compressed(raw[0])
sc = random_fr_limbs(4096, 7)
kzg.lincomb_g1(pts, sc)
t0 = time.perf_counter()
out = kzg.lincomb_g1(pts, sc)
print("lincomb_g1 n=4096 host-buffer call: %.3f ms" % ((time.perf_counter() - t0) * 1e3))
