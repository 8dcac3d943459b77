#!/usr/bin/env python
"""One-polynomial FK20Single (n = 4096) from 1..32 host threads, in both latency modes (b200_set_latency_mode): the
steady-state aggregate rate.  Run from the repository root on a B200: PYTHONPATH=. python tools/conc_probe.py"""
import time, numpy as np, threading
import go_kzg_b200 as kzg
from go_kzg_b200.synth import random_fr_limbs
raw = np.fromfile("tests/golden/trusted_setup_g1.bin", dtype=np.uint8).reshape(2, 4096, 48)
first = kzg.g1_from_compressed(raw[0])
rest = kzg.g1_mul_many(np.repeat(first[:1], 4096, axis=0), kzg.fr_from_ints([pow(1337, i, kzg.R_MOD) for i in range(4096, 8192)]))
fs = kzg.FFTSettings(13)
ks = kzg.KZGSettings(fs, np.concatenate([first, rest]))
fk = kzg.FK20SingleSettings(ks, 8192)
poly = random_fr_limbs(4096, 1)
fk.fk20_single(poly)
def conc(nt, reps):
    ps = [random_fr_limbs(4096, 100 + t) for t in range(nt)]
    def w(t):
        for _ in range(reps): fk.fk20_single(ps[t])
    th = [threading.Thread(target=w, args=(t,)) for t in range(nt)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
    return nt * reps / (time.perf_counter() - t0)
for rnd in range(2):
  for mode in (1, 0):
    kzg.lib().b200_set_latency_mode(mode)
    conc(4, 1)
    print("mode %d, 10 calls per thread, polynomials/s:" % mode, {t: round(conc(t, 10), 1) for t in (1, 2, 3, 4, 8, 16, 32)})
