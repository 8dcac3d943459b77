#!/usr/bin/env python
"""Per-kernel measurements for the rows of SURVEY.md 8d that are not the headline metric:
Fr NTT, DAS extension, recovery, G1 FFT, MSM -- device time per unit (library CUDA events around
every launch), achieved algorithmic GB/s against the measured HBM peak.  Run on a B200:
    python tools/bench_components.py > profiles/rNN_components.json
Host-buffer entry points are used, so the event-timed kernel classes (not the copies) are reported."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go_kzg_b200 as kzg                       # noqa: E402
from go_kzg_b200.synth import random_fr_limbs   # noqa: E402

L = kzg.lib()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
CLASSES = ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "misc"]


def timed(fn, reps=3):
    fn()                                         # warm-up (tables, pools)
    best = None
    for _ in range(reps):
        L.b200_profile_begin()
        fn()
        ms = (C.c_double * 5)()
        n = (C.c_uint64 * 5)()
        L.b200_profile_end(ms, n)
        cur = {k: ms[i] for i, k in enumerate(CLASSES)}
        cur["launches"] = int(sum(n))
        if best is None or sum(cur[k] for k in CLASSES) < sum(best[k] for k in CLASSES):
            best = cur
    return best


def row(name, units, algo_bytes_per_unit, t, cls):
    ms = sum(t[k] for k in cls)
    gbs = units * algo_bytes_per_unit / (ms / 1e3) / 1e9
    return {"kernel": name, "units": units, "algo_bytes_per_unit": algo_bytes_per_unit, "device_ms": round(ms, 4),
            "units_per_s": round(units / (ms / 1e3), 1), "achieved_GBs": round(gbs, 2), "hbm_peak_GBs": PEAK,
            "frac_of_hbm": round(gbs / PEAK, 5), "launches": t["launches"], "classes_ms": {k: round(t[k], 4) for k in CLASSES}}


def main():
    out = []
    for scale, batch in ((10, 4096), (12, 2048), (13, 1024), (14, 512)):
        n = 1 << scale
        fs = kzg.FFTSettings(scale)
        v = np.stack([random_fr_limbs(n, 1)] * batch)
        t = timed(lambda: fs.fft_batch(v))
        out.append(row("fr_ntt n=2^%d batch=%d" % (scale, batch), batch, 64 * n, t, ["fr_ntt"]))
    fs = kzg.FFTSettings(14)
    n, batch = 8192, 512
    v = np.stack([random_fr_limbs(n, 2)] * batch)
    t = timed(lambda: fs.das_fft_extension_batch(v))
    out.append(row("das_fft_extension n=8192 (scale 14) batch=%d" % batch, batch, 2 * 32 * n, t, ["fr_ntt"]))
    # recovery, config 4: n = 2^14, 50 % missing
    n, batch = 1 << 14, 64
    even = random_fr_limbs(n // 2, 3)
    odd = fs.das_fft_extension(even)
    full = np.empty((n, 4), dtype=np.uint64)
    full[0::2], full[1::2] = even, odd
    rng = np.random.default_rng(14)
    present = np.ones((batch, n), dtype=np.uint8)
    for b in range(batch):
        present[b, rng.permutation(n)[: n // 2]] = 0
    samples = np.stack([full] * batch)
    rec = None

    def do_rec():
        nonlocal rec
        rec = fs.recover_poly_from_samples_batch(samples, present)
    t = timed(do_rec)
    assert np.array_equal(rec[0], full) and np.array_equal(rec[-1], full)
    out.append(row("recover_poly_from_samples n=2^14 50%% missing batch=%d" % batch, batch, 1_064_960, t, ["fr_ntt", "misc"]))
    # G1 FFT 4096 (single transform and batch 32) and MSM 4096 over the trusted setup
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "trusted_setup_g1.bin"), dtype=np.uint8).reshape(2, 4096, 48)
    pts = kzg.g1_from_compressed(raw[0])
    fs12 = kzg.FFTSettings(12)
    t = timed(lambda: fs12.fft_g1(pts, True), reps=2)
    out.append(row("fft_g1 n=4096 inverse, single transform", 1, 288 * 4096, t, ["g1_fft_stage", "g1_mul"]))
    pb = np.stack([pts] * 32)
    t = timed(lambda: fs12.fft_g1_batch(pb, False), reps=2)
    out.append(row("fft_g1 n=4096 forward, batch=32", 32, 288 * 4096, t, ["g1_fft_stage", "g1_mul"]))
    ks = kzg.KZGSettings(fs12, pts)
    co = np.stack([random_fr_limbs(4096, 100 + b) for b in range(256)])
    t = timed(lambda: ks.commit_to_poly_batch(co))
    out.append(row("commit_to_poly n=4096 (fixed-base tables) batch=256", 256, 721_040, t, ["g1_mul", "g1_fold"]))
    sc = random_fr_limbs(4096, 7)
    t = timed(lambda: kzg.lincomb_g1(pts, sc))
    out.append(row("lincomb_g1 n=4096 generic (no tables), single", 1, 721_040, t, ["g1_mul", "g1_fold"]))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
