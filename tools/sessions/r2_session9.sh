#!/bin/bash
# Round 2, GPU session 9: quad operations without divergent operand selection -- MSM tails and the quad stage kernel
# (variant library with B200_QUAD_STAGE_MAX=4096) against the default.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "bucket_msm or lane_mappings or commit_without_table" 2>&1 | tail -2
for lib in "" "go_kzg_b200/lib/libb200kzg_quadstage.so"; do
  echo "== lib: ${lib:-default}"
  B200_KZG_LIB=$lib timeout 300 python tools/msm_probe.py 2>&1 | tail -2
  B200_KZG_LIB=$lib timeout 600 python - <<'PY'
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
ref = fk.fk20_single(poly)
def wall(fn, reps=3):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e3
print("fk20_single one polynomial: %.2f ms" % wall(lambda: fk.fk20_single(poly)))
print("da_using_fk20 one polynomial: %.2f ms" % wall(lambda: fk.da_using_fk20(poly)))
print("fft_g1 n=4096: %.2f ms" % wall(lambda: fs.fft_g1(first)))
def conc(nt, reps=2):
    ps = [random_fr_limbs(4096, 100 + t) for t in range(nt)]
    def w(t):
        for _ in range(reps): fk.fk20_single(ps[t])
    th = [threading.Thread(target=w, args=(t,)) for t in range(nt)]
    t0 = time.perf_counter(); [x.start() for x in th]; [x.join() for x in th]
    return nt * reps / (time.perf_counter() - t0)
conc(4, 1)
print("concurrent callers polys/s:", {t: round(conc(t), 1) for t in (1, 8, 32)})
import hashlib
print("proof digest", hashlib.sha256(kzg.g1_to_compressed(ref).tobytes()).hexdigest()[:16])
PY
done
