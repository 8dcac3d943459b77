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
CLASSES = ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "misc", "g1_lookup", "g1_msm"]


def timed(fn, reps=3):
    fn()                                         # warm-up (tables, pools)
    best = None
    for _ in range(reps):
        L.b200_profile_begin()
        fn()
        ms = (C.c_double * len(CLASSES))()
        n = (C.c_uint64 * len(CLASSES))()
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
    out.append(row("commit_to_poly n=4096 (fixed-base tables) batch=256", 256, 721_040, t, ["g1_lookup", "g1_fold"]))
    sc = random_fr_limbs(4096, 7)
    t = timed(lambda: kzg.lincomb_g1(pts, sc))
    out.append(row("lincomb_g1 n=4096 generic (bucket MSM, affine points), single", 1, 721_040, t, ["g1_msm"]))
    # Jacobian inputs (Z != 1): general additions in the buckets
    gen = np.zeros((1, 18), dtype=np.uint64)
    L.b200_g1_generator(gen.ctypes.data)
    jac = kzg.g1_mul_many(np.repeat(gen, 4096, axis=0), random_fr_limbs(4096, 8))
    t = timed(lambda: kzg.lincomb_g1(jac, sc))
    out.append(row("lincomb_g1 n=4096 generic (bucket MSM, Jacobian points), single", 1, 721_040, t, ["g1_msm"]))
    big_n = 1 << 16
    big = np.concatenate([pts] * (big_n // 4096))
    bsc = random_fr_limbs(big_n, 9)
    t = timed(lambda: kzg.lincomb_g1(big, bsc), reps=2)
    r_msm = row("lincomb_g1 n=65536 generic (bucket MSM), single", 1, big_n * 176 + 144, t, ["g1_msm"])
    r_msm["terms_per_s"] = round(big_n / (r_msm["device_ms"] / 1e3), 1)
    out.append(r_msm)
    t = timed(lambda: kzg.g1_mul_many(big, bsc), reps=2)
    r_var = row("g1_mul_many n=65536 (k_g1_mul_var: per-term windowed GLV multiplication)", 1, big_n * 320, t, ["g1_mul"])
    r_var["terms_per_s"] = round(big_n / (r_var["device_ms"] / 1e3), 1)
    out.append(r_var)
    # where the fixed tail of the bucket method (bucket reduction + 120 doublings, ~1.1 ms) stops mattering
    for lg in (18, 20):
        nn = 1 << lg
        pp = np.concatenate([pts] * (nn // 4096))
        ss = random_fr_limbs(nn, 20 + lg)
        t = timed(lambda: kzg.lincomb_g1(pp, ss), reps=2)
        r = row("lincomb_g1 n=2^%d generic (bucket MSM), single" % lg, 1, nn * 176 + 144, t, ["g1_msm"])
        r["terms_per_s"] = round(nn / (r["device_ms"] / 1e3), 1)
        r["per_term_rate_vs_k_g1_mul_var"] = round(r["terms_per_s"] / r_var["terms_per_s"], 2)
        out.append(r)
        del pp, ss
    r_msm["per_term_rate_vs_k_g1_mul_var"] = round(r_msm["terms_per_s"] / r_var["terms_per_s"], 2)
    # one polynomial per call (the reference API): FK20Single / DAUsingFK20 latency at n = 4096
    import time
    first = pts
    from go_kzg_b200.kzg import R_MOD
    rest = kzg.g1_mul_many(np.repeat(gen, 4096, axis=0), kzg.fr_from_ints([pow(1337, i, R_MOD) for i in range(4096, 8192)]))
    fs13 = kzg.FFTSettings(13)
    fk = kzg.FK20SingleSettings(kzg.KZGSettings(fs13, np.concatenate([first, rest])), 8192)
    poly = random_fr_limbs(4096, 11)
    for name, fn in (("fk20_single n=4096, one polynomial per call", lambda: fk.fk20_single(poly)),
                     ("da_using_fk20 n=4096, one polynomial per call", lambda: fk.da_using_fk20(poly))):
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        wall = (time.perf_counter() - t0) / 3
        t = timed(fn, reps=2)
        r1 = row(name, 1, 1_900_544, t, ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "g1_lookup", "misc"])
        r1["wall_ms_per_call"] = round(wall * 1e3, 3)
        out.append(r1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
