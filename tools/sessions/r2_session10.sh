#!/bin/bash
# Round 2, GPU session 10: adaptive quad stage kernels (default build) -- full GPU suite, one-polynomial latencies,
# concurrent callers, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/quad_stage.txt
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
def wall(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
def conc(nt, reps=2):
    ps = [random_fr_limbs(4096, 100 + t) for t in range(nt)]
    def w(t):
        for _ in range(reps): fk.fk20_single(ps[t])
    th = [threading.Thread(target=w, args=(t,)) for t in range(nt)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
    return nt * reps / (time.perf_counter() - t0)
for mode in (1, 0):
    kzg.lib().b200_set_latency_mode(mode)
    print("latency mode %d: FK20Single %.2f ms, DAUsingFK20 %.2f ms, FFTG1(4096) %.2f ms, FFTG1(8192) %.2f ms" % (
        mode, wall(lambda: fk.fk20_single(poly)), wall(lambda: fk.da_using_fk20(poly)), wall(lambda: fs.fft_g1(first)),
        wall(lambda: fs.fft_g1(np.concatenate([first, rest])))))
    conc(4, 1)
    print("   concurrent callers, polynomials/s:", {t: round(conc(t), 1) for t in (1, 2, 4, 8, 32)})
kzg.lib().b200_set_latency_mode(1)
PY
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
