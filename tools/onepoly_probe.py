#!/usr/bin/env python
"""Per-class device time of ONE FK20Single / DAUsingFK20 call (n = 4096) -- where the one-polynomial latency goes.
PYTHONPATH=. python tools/onepoly_probe.py"""
import ctypes as C
import time

import numpy as np

import go_kzg_b200 as kzg
from go_kzg_b200.synth import random_fr_limbs

L = kzg.lib()
raw = np.fromfile("tests/golden/trusted_setup_g1.bin", dtype=np.uint8).reshape(2, 4096, 48)
first = kzg.g1_from_compressed(raw[0])
rest = kzg.g1_mul_many(np.repeat(first[:1], 4096, axis=0), kzg.fr_from_ints([pow(1337, i, kzg.R_MOD) for i in range(4096, 8192)]))
fs = kzg.FFTSettings(13)
ks = kzg.KZGSettings(fs, np.concatenate([first, rest]))
fk = kzg.FK20SingleSettings(ks, 8192)
poly = random_fr_limbs(4096, 1)
names = ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "misc", "g1_lookup", "g1_msm"]
for mode in (1, 0):
    L.b200_set_latency_mode(mode)
    for label, fn in (("FK20Single", lambda: fk.fk20_single(poly)), ("DAUsingFK20", lambda: fk.da_using_fk20(poly)),
                      ("CommitToPoly", lambda: ks.commit_to_poly(poly)), ("commit+FK20 batch of 1", lambda: fk.commit_fk20_batch(poly.reshape(1, 4096, 4)))):
        fn()
        ms = (C.c_double * 7)()
        cnt = (C.c_uint64 * 7)()
        L.b200_profile_begin()
        t0 = time.perf_counter()
        fn()
        wall = (time.perf_counter() - t0) * 1e3
        L.b200_profile_end(ms, cnt)
        print("mode %d %-24s wall %6.2f ms | " % (mode, label, wall) + "  ".join("%s %.2f/%d" % (n, ms[i], cnt[i]) for i, n in enumerate(names) if cnt[i]))
L.b200_set_latency_mode(1)
