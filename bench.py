#!/usr/bin/env python
"""Headline benchmark: blobs/s for CommitToPoly + FK20Single all-proofs at n = 4096
(BASELINE.json metric), on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --steps K --warmup W    # reference algorithm on the host CPUs

One "step" = one batch of `--batch` synthetic blobs per GPU through the hot path
(b200_commit_fk20_batch_dev: Fr NTT of the Toeplitz coefficients, fixed-setup scalar
multiplications, G1 inverse + forward FFT, MSM commitment).  `value` is timed with inputs
resident in HBM (CUDA events on the launching stream); `e2e` goes through the host-buffer C-ABI
call (b200_commit_fk20_batch: H2D of the polynomials, D2H of commitments and proofs inside the
timed region).  Multi-GPU = blob-parallel replicas (weak scaling, no data-path collective);
torch.distributed over NCCL is used only for the barrier and the max-over-ranks of the time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_COEFFS = int(os.environ.get("B200_BENCH_N", "4096"))   # 4096 is the metric's size; the override exists for the CPU contract test
SCALE = N_COEFFS.bit_length()                             # FK20 over n coefficients needs the 2n domain
SECRET = 1337                       # eth/trusted_setup.json's (insecure, known) secret
ALGO_BYTES_PER_BLOB = 2_621_584     # commit (721 040) + FK20Single (1 900 544)
G1_NTT_BYTES = lambda n: 288 * n    # one G1 NTT of n points
PROFILE_CLASSES = ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "misc", "g1_lookup", "g1_msm"]   # include/b200_kzg.h B200_PROFILE_CLASSES
N_CLASSES = len(PROFILE_CLASSES)
METRIC = "blobs/sec (commit+FK20 all-proofs, n=%d)" % N_COEFFS


def load_json(rel):
    try:
        with open(os.path.join(ROOT, rel)) as f:
            return json.load(f)
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        # median over the busiest half of the samples (the region under load)
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


def setup_points(kzg):
    """SecretG1 = [1337^i G], i < 8192: first 4096 from the reference's fixture bytes, the rest
    generated on the device (same values; test_gpu_parity checks them against the oracle)."""
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "trusted_setup_g1.bin"), dtype=np.uint8).reshape(2, 4096, 48)
    first = kzg.g1_from_compressed(raw[0])
    R = kzg.R_MOD
    gen = np.zeros((1, 18), dtype=np.uint64)
    kzg.lib().b200_g1_generator(gen.ctypes.data)
    scal = kzg.fr_from_ints([pow(SECRET, i, R) for i in range(4096, 8192)])
    rest = kzg.g1_mul_many(np.repeat(gen, 4096, axis=0), scal)
    return np.concatenate([first, rest])


IMAD_PER_MUL, IMAD_PER_SQR = 300, 234   # 32x32->64 multiply-adds of fe_mul (2 N^2 + N) and fe_sqr (N (N + 1) / 2 + N^2 + N), N = 12


def fp_ops_model_per_blob():
    """(Fp products, Fp squares) one blob needs on our path (DESIGN.md 'integer roofline'):
    fixed-base products for ToeplitzPart2 (2n) and the commitment (n), two size-n G1 transforms
    with width-5 NAF GLV twiddle programs, n twists, n final additions, the commitment fold."""
    dbl, add, mixed = (2, 5), (12, 4), (8, 3)                               # (M, S) of the group operations
    def cost(n_dbl=0, n_add=0, n_mixed=0, m=0, s=0):
        return (n_dbl * dbl[0] + n_add * add[0] + n_mixed * mixed[0] + m, n_dbl * dbl[1] + n_add * add[1] + n_mixed * mixed[1] + s)
    def plus(*cs):
        return (sum(c[0] for c in cs), sum(c[1] for c in cs))
    def times(k, c):
        return (k * c[0], k * c[1])
    # fixed-twiddle program (GLV, width-5 NAF): 128 doublings, ~43 mixed additions against an
    # effective-affine table (1 doubling, 7 mixed additions, 4 + 34 rescaling products of which 7 are
    # squares), 8 beta products, 1 product to leave the isomorphic curve
    wnaf5 = cost(n_dbl=129, n_mixed=2 * 128 / 6 + 7, m=3 + 27 + 8 + 1, s=1 + 7)
    fixed_base = cost(n_mixed=22)                                           # 22 signed 12-bit windows
    butterfly = (14, 4)                                                     # shared-subexpression add/sub pair
    n = N_COEFFS
    stages = (n // 2) * (n.bit_length() - 1)
    fft = plus(times(stages - (n - 1), wnaf5), times(stages, butterfly))    # trivial twiddles skip the product
    # the inverse transform loses its first two stages (folded into ToeplitzPart2 as 4-term look-up sums):
    # n butterflies and n - 3 non-trivial twiddle products less, 3 n more look-up products
    fft_inv = plus(times(stages - (n - 1) - (n - 3), wnaf5), times(stages - n, butterfly))
    return plus(fft, fft_inv, times(n, wnaf5), times(n, cost(n_add=1)), times(3 * n + 3 * n, fixed_base), times(n - 1, cost(n_add=1)))


def run_components(kzg, L, fk, fs, torch, dist, rank, world, max_over_ranks, barrier):
    """Numbers the headline metric does not show, measured at every N (device time with CUDA events where the work is
    device-resident, wall clock through the host-buffer call otherwise; max over ranks):
      * one polynomial per call (the reference's API shape): FK20Single and DAUsingFK20 at n = 4096;
      * config 4: DASFFTExtension + RecoverPolyFromSamples, n = 2^14, half missing, `batch` polynomials per GPU;
      * config 5: DAUsingFK20Multi n = 2^20, chunk 16, chunk offsets sharded over the N GPUs with one exchange of
        partial G1 sums over NCCL (go_kzg_b200/multi_gpu.py)."""
    from go_kzg_b200.synth import random_fr_limbs
    from go_kzg_b200 import multi_gpu
    out = {}

    def wall(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return max_over_ranks((time.perf_counter() - t0) / reps)

    poly = random_fr_limbs(N_COEFFS, 0xC0DE + rank)
    barrier()
    out["fk20_single_one_polynomial_ms"] = round(wall(lambda: fk.fk20_single(poly), 3) * 1e3, 3)
    out["da_using_fk20_one_polynomial_ms"] = round(wall(lambda: fk.da_using_fk20(poly), 3) * 1e3, 3)
    if N_COEFFS == 4096:
        dab = 32
        dpolys = np.stack([random_fr_limbs(N_COEFFS, 0xDA00 + rank * dab + b) for b in range(dab)])
        t_da = wall(lambda: fk.da_using_fk20_batch(dpolys), 2)
        out["da_using_fk20_batch"] = {"blobs_per_s": round(world * dab / t_da, 2), "batch_per_gpu": dab,
                                      "note": "b200_da_using_fk20_batch, n = 4096 -> 8192 proofs per blob, host buffers (copies included), wall clock"}
    # the same one-polynomial call from many host threads at once (goroutines in the Go binding): every call runs on its own
    # pooled stream, so independent callers fill the SMs that a single 64-warp stage leaves idle
    import threading
    def concurrent(nthreads, reps):
        polys_t = [random_fr_limbs(N_COEFFS, 0xCC00 + rank * 64 + t) for t in range(nthreads)]
        errs = []
        def worker(t):
            try:
                for _ in range(reps):
                    fk.fk20_single(polys_t[t])
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        dt = time.perf_counter() - t0
        if errs:
            raise errs[0]
        return nthreads * reps / dt
    concurrent(32, 1)          # creates the pooled streams and grows the memory pool outside the timed passes
    out["fk20_single_concurrent_callers"] = {("%d_threads_polys_per_s" % t): round(world * max_over_ranks(-concurrent(t, 6)) * -1, 2) for t in (1, 2, 8, 32)}
    out["one_polynomial_note"] = "host-buffer call per polynomial, n = %d (b200_fk20_single / b200_da_using_fk20), wall clock, max over ranks" % N_COEFFS
    if N_COEFFS == 4096:
        # ---- verification: all 4096 FK20Single proofs of one blob against its commitment.  Aggregated: random linear
        # combination, three device MSMs + ONE host pairing; per proof: the reference's CheckProofSingle, pairing per proof
        # spread over the host cores (a 64-proof sample).  Both must say "valid".
        s2 = kzg.generate_testing_setup_g2(SECRET, 2)
        rc = L.b200_kzg_settings_set_secret_g2(fk.ks.h, s2.ctypes.data, 2)
        assert rc == 0, L.b200_strerror(rc)
        commits, proofs = fk.commit_fk20_batch(poly.reshape(1, N_COEFFS, 4))
        w = int(kzg.fr_to_ints(kzg.FFTSettings(12).expanded_roots_of_unity()[1:2])[0])
        xs_i, acc = [], 1
        for _ in range(N_COEFFS):
            xs_i.append(acc)
            acc = acc * w % kzg.R_MOD
        xs = kzg.fr_from_ints(xs_i)
        ys = kzg.FFTSettings(12).fft(poly)
        cs = np.repeat(commits[0].reshape(1, 18), N_COEFFS, axis=0)
        agg_ok = []
        rs = fk.ks._random_scalars(N_COEFFS)      # the caller's CSPRNG draw (crypto/rand in the Go shim), outside the timed region
        t_agg = wall(lambda: agg_ok.append(fk.ks.check_proof_single_aggregate(cs, proofs[0], xs, ys, rs=rs)), 3)
        one_ok = []
        t_one = wall(lambda: one_ok.append(bool(fk.ks.check_proof_single_batch(cs[:64], proofs[0][:64], xs[:64], ys[:64]).all())), 1)
        out["verification"] = {"aggregate_4096_proofs_ms": round(t_agg * 1e3, 2), "aggregate_proofs_per_s": round(world * N_COEFFS / t_agg, 1),
                               "per_proof_checks_per_s": round(64 / t_one, 1), "host_cores": os.cpu_count(), "all_valid": bool(all(agg_ok) and all(one_ok)),
                               "note": "b200_check_proof_single_aggregate (3 bucket MSMs of 4096 terms on the device + 1 pairing on the host) vs "
                                       "b200_check_proof_single_batch (one host pairing check per proof, all host cores); host buffers, wall clock"}
        # ---- config 4
        scale, batch = 14, 64
        n = 1 << scale
        fs14 = kzg.FFTSettings(scale)
        even = np.stack([random_fr_limbs(n // 2, 0xD4000000 + rank * batch + b) for b in range(batch)])
        odd = fs14.das_fft_extension_batch(even)
        full = np.empty((batch, n, 4), dtype=np.uint64)
        full[:, 0::2], full[:, 1::2] = even, odd
        rng = np.random.default_rng(14 + rank)
        present = np.ones((batch, n), dtype=np.uint8)
        for b in range(batch):
            present[b, rng.permutation(n)[: n // 2]] = 0
        # pinned host buffers, raw C-ABI calls (the numpy convenience wrappers allocate pageable outputs)
        def pinned(a):
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64 if a.dtype == np.uint64 else a.dtype)).pin_memory()
            return t, t.numpy().view(a.dtype)
        full_ref = full.copy()
        full[present == 0] = 0
        t_s, samples = pinned(full)
        t_p, pres = pinned(present)
        t_o, rec = pinned(np.zeros_like(full))
        t_e, ev_io = pinned(even)
        def do_ext():   # in place: the timed repetitions extend the previous output again (same work, different values)
            rc = L.b200_das_fft_extension_batch(fs14.h, ev_io.ctypes.data, n // 2, batch)
            assert rc == 0, L.b200_strerror(rc)
        do_ext()
        ext_ok = bool(np.array_equal(ev_io, odd))
        def do_rec():
            rc = L.b200_recover_poly_from_samples_batch(fs14.h, samples.ctypes.data, pres.ctypes.data, n, batch, rec.ctypes.data)
            assert rc == 0, L.b200_strerror(rc)
        barrier()
        t_ext = wall(do_ext, 3)
        t_rec = wall(do_rec, 3)
        ok = bool(np.array_equal(rec, full_ref)) and ext_ok
        out["config4"] = {"workload": "DASFFTExtension(8192 -> 16384) + RecoverPolyFromSamples(n=2^14, 50%% missing), %d polynomials per GPU per call" % batch,
                          "das_extension_polys_per_s": round(world * batch / t_ext, 1), "recover_polys_per_s": round(world * batch / t_rec, 1),
                          "recovered_equals_original": ok,
                          "h2d_bytes_per_call": int(samples.nbytes + pres.nbytes), "d2h_bytes_per_call": int(rec.nbytes),
                          "timing": "b200_das_fft_extension_batch / b200_recover_poly_from_samples_batch with pinned host buffers (copies included), wall clock, max over ranks"}
        del fs14
    return out


def run_config5(args, dist, world):
    """config 5 measured after the main handles are released (its tables take up to 103 GB on one GPU)"""
    from go_kzg_b200 import multi_gpu
    return multi_gpu.measure_da_using_fk20_multi_sharded(scale=args.config5_scale, chunk_len=16, reps=2, dist=dist if world > 1 else None)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import go_kzg_b200 as kzg
    from go_kzg_b200 import shard
    from go_kzg_b200.synth import blob_polys

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = kzg.lib()
    assert L.b200_set_device(local) == 0
    B, K, W = args.batch, args.steps, args.warmup

    fs = kzg.FFTSettings(SCALE)
    fk = kzg.FK20SingleSettings(kzg.KZGSettings(fs, setup_points(kzg)), 2 * N_COEFFS)
    polys = blob_polys(B, N_COEFFS, first_blob=shard.blob_range(rank, world, B).start)   # (B, 4096, 4) u64
    h_polys = torch.from_numpy(polys.view(np.int64)).pin_memory()
    d_polys = h_polys.cuda(non_blocking=True)
    d_commit = torch.zeros((B, 18), dtype=torch.int64, device="cuda")
    d_proofs = torch.zeros((B, N_COEFFS, 18), dtype=torch.int64, device="cuda")
    h_commit = torch.zeros((B, 18), dtype=torch.int64).pin_memory()
    h_proofs = torch.zeros((B, N_COEFFS, 18), dtype=torch.int64).pin_memory()
    h_commit48 = torch.zeros((B, 48), dtype=torch.uint8).pin_memory()
    h_proofs48 = torch.zeros((B, N_COEFFS, 48), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")   # > 126 MB L2
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)

    def step_dev():
        flush.zero_()
        rc = L.b200_commit_fk20_batch_dev(fk.h, d_polys.data_ptr(), N_COEFFS, B, d_commit.data_ptr(), d_proofs.data_ptr(), sptr)
        if rc:
            raise RuntimeError("b200_commit_fk20_batch_dev: %s / %s" % (L.b200_strerror(rc), L.b200_last_cuda_error()))

    def step_e2e():
        rc = L.b200_commit_fk20_batch(fk.h, h_polys.data_ptr(), N_COEFFS, B, h_commit.data_ptr(), h_proofs.data_ptr())
        if rc:
            raise RuntimeError("b200_commit_fk20_batch: %s / %s" % (L.b200_strerror(rc), L.b200_last_cuda_error()))

    def step_e2e_compressed():
        rc = L.b200_commit_fk20_batch_compressed(fk.h, h_polys.data_ptr(), N_COEFFS, B, h_commit48.data_ptr(), h_proofs48.data_ptr())
        if rc:
            raise RuntimeError("b200_commit_fk20_batch_compressed: %s / %s" % (L.b200_strerror(rc), L.b200_last_cuda_error()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return shard.max_over_ranks(x, dist if world > 1 else None, device="cuda")

    # ---- device-resident timing (value) + per-kernel-class events (roofline)
    for _ in range(W):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    L.b200_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        step_dev()
    e1.record(stream)
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    cls_ms = (C.c_double * N_CLASSES)()
    cls_n = (C.c_uint64 * N_CLASSES)()
    L.b200_profile_end(cls_ms, cls_n)
    launches = fk.last_launch_count() * K
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer ABI call
    for _ in range(max(1, min(W, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    # same call with 48-byte compressed outputs produced on the device (what the callers of the path consume)
    step_e2e_compressed()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        step_e2e_compressed()
    torch.cuda.synchronize()
    e2e_c_s = max_over_ranks(time.perf_counter() - t0)
    barrier()

    # Closed-form spot check of one blob of the timed batch, outside the timed region: with the setup's known
    # secret s the commitment is p(s) G and proof i is ((p(s) - p(w^i)) / (s - w^i)) G (fk20_single.go:122-134).
    # The expected points come from Python integers and the library's host-side double-and-add (b200_g1_mul),
    # which shares no code path with the device pipeline; compared projectively (bls/bls_kilic.go:106 EqualG1).
    def closed_form_ok(commit_row, proof_rows, poly_limbs):
        R = kzg.R_MOD
        coeffs = kzg.fr_to_ints(poly_limbs)
        def horner(x):
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % R
            return acc
        gen = np.zeros(18, dtype=np.uint64)
        L.b200_g1_generator(gen.ctypes.data)
        def expect(k):
            out = np.zeros(18, dtype=np.uint64)
            kk = kzg.fr_from_ints([k % R])
            L.b200_g1_mul(out.ctypes.data, gen.ctypes.data, kk.ctypes.data)
            return out
        ps = horner(SECRET)
        got = np.ascontiguousarray(commit_row, dtype=np.uint64)
        ok = L.b200_g1_equal(got.ctypes.data, expect(ps).ctypes.data) == 1
        w = pow(7, (R - 1) // N_COEFFS, R)                      # bls/globals.go: primitive root 7 -> root of order n
        for i in (0, 1, N_COEFFS // 2 - 1, N_COEFFS - 1):
            x = pow(w, i, R)
            q = (ps - horner(x)) * pow((SECRET - x) % R, -1, R) % R
            got = np.ascontiguousarray(proof_rows[i], dtype=np.uint64)
            ok = ok and L.b200_g1_equal(got.ctypes.data, expect(q).ctypes.data) == 1
        return bool(ok)
    chk = B - 1
    same = closed_form_ok(d_commit[chk].cpu().numpy().view(np.uint64), d_proofs[chk].cpu().numpy().view(np.uint64), polys[chk]) and \
        closed_form_ok(h_commit[0].numpy().view(np.uint64), h_proofs[0].numpy().view(np.uint64), polys[0])

    components = None
    if not args.no_components:
        components = run_components(kzg, L, fk, fs, torch, dist, rank, world, max_over_ranks, barrier)
        if args.config5_scale > 0 and N_COEFFS == 4096:
            fk.close(); fk.ks.close()
            del d_proofs, h_proofs, h_proofs48, flush
            torch.cuda.empty_cache()
            components["config5"] = run_config5(args, dist, world)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm_peak, peak_kind = measured_peaks()
    value = shard.whole_job_rate(B, world, K, ms_total / 1e3)
    # dominant kernel: the G1 FFT butterfly stage.  One launch = one stage of B size-n transforms;
    # a transform's algorithmic bytes (288 n) are apportioned over its log2(n) stages
    # and the step runs two size-n transforms per blob.
    stage_ms, stage_n = cls_ms[1], max(1, cls_n[1])
    algo_stage_bytes = B * 2 * G1_NTT_BYTES(N_COEFFS) * K / stage_n                       # average per launch
    achieved = algo_stage_bytes / (stage_ms / stage_n / 1e3) / 1e9
    tj = load_json("profiles/stage_kernel_traffic.json")   # per-launch DRAM bytes / instruction counts from the committed ncu --set full capture
    traffic = tj["dram_bytes_per_launch"] * (B / tj["blobs_per_launch"]) if tj else None
    # Integer roofline of the dominant kernel, MEASURED: lane-level IMAD.WIDE operations one launch executes (ncu:
    # inst_executed x share of IMAD.WIDE among executed instructions x 32, profiles/stage_kernel_traffic.json, written by
    # tools/ncu_record.py from the committed capture) over the launch duration timed live above, against the issue limit
    # of the multiplier pipe: one warp-wide IMAD.WIDE per 4 cycles per SM sub-partition = 8 lane-operations per cycle.
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    imad_peak = 148 * 4 * 8 * sm_mhz * 1e6
    imad_meas = None
    if tj and tj.get("imad_wide_lane_ops_per_launch"):
        imad_meas = tj["imad_wide_lane_ops_per_launch"] * (B / tj["blobs_per_launch"]) / (stage_ms / stage_n / 1e3)
    # the model count kept beside it: Fp products and squares per blob x multiply-adds each, against the multiplier probe
    pm = C.c_float()
    threads = 148 * 2048
    L.b200_probe_fp_mul(threads, 2000, C.byref(pm))
    fp_peak = threads * 2000 * 2 / (pm.value / 1e3)
    n_mul, n_sqr = fp_ops_model_per_blob()
    imad_model = (n_mul * IMAD_PER_MUL + n_sqr * IMAD_PER_SQR) * value / world
    # per kernel class: device time per step, algorithmic bytes per step (arrays in + out, SURVEY.md 8d sizes), GB/s
    n = N_COEFFS
    class_bytes = {
        "fr_ntt": B * 64 * 2 * n,                                             # one Fr NTT of 2n points per blob
        "g1_fft_stage": B * 2 * G1_NTT_BYTES(n),                              # two size-n G1 transforms per blob
        "g1_mul": B * 288 * n,                                                # the twist between the two transforms (in place)
        "g1_fold": B * (n * 144 + 144) + B * 3 * 144 * n,                     # commitment fold + final addition of the even slots
        "g1_lookup": B * (2 * n * (32 + 144) + n * (32 + 144)) + 3 * n * 144,  # ToeplitzPart2 (2n) + commitment terms (n), bases once
        "misc": B * (n * 32 + 2 * n * 32 + n * 144 + n * 144),
    }
    kernels = {}
    for i, k in enumerate(PROFILE_CLASSES):
        if cls_n[i] == 0:
            continue
        ms = cls_ms[i] / K
        ab = class_bytes.get(k)
        kernels[k] = {"ms_per_step": round(ms, 3), "launches_per_step": int(cls_n[i] // K), "share_of_step": round(cls_ms[i] / ms_total, 4)}
        if ab:
            gbs = ab / (ms / 1e3) / 1e9
            kernels[k].update({"algo_bytes_per_step": int(ab), "achieved_GBs": round(gbs, 3), "frac_of_hbm": round(gbs / hbm_peak, 6)})
    for k, rec in (load_json("profiles/kernel_pipe_records.json") or {}).items():   # pipe utilisation per kernel from committed ncu captures
        if k in kernels:
            kernels[k]["ncu"] = rec
    out = {
        "metric": METRIC, "value": round(value, 3), "unit": "blobs/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"arithmetic": "381-bit Fp / 255-bit Fr Montgomery integers on 32-bit limbs (IMAD.WIDE)", "workload": "CommitToPoly + FK20Single, n=4096 (configs[1]+configs[2]), eth trusted setup secret 1337 "
                               "extended to 8192 points", "blobs_per_gpu_per_step": B, "l2": "256 MiB flush write between steps",
                   "parallelism": "blob-parallel replicas x%d" % world, "closed_form_spot_check": same},
        "e2e": {"value": round(world * B * K / e2e_s, 3), "unit": "blobs/s",
                "h2d_bytes_per_step": int(B * N_COEFFS * 32), "d2h_bytes_per_step": int(B * (N_COEFFS + 1) * 144)},
        "e2e_compressed": {"value": round(world * B * K / e2e_c_s, 3), "unit": "blobs/s", "h2d_bytes_per_step": int(B * N_COEFFS * 32),
                           "d2h_bytes_per_step": int(B * (N_COEFFS + 1) * 48),
                           "note": "b200_commit_fk20_batch_compressed: ToCompressedG1 on the device, 48 B per point back to the host"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_g1_fft_stage", "achieved": round(achieved, 4), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 6), "traffic": traffic, "peak_kind": peak_kind,
                     "share_of_step": round(stage_ms / ms_total, 4), "launches": int(stage_n),
                     "note": "integer-pipe bound kernel; see int_roofline"},
        "int_roofline": {"kernel": "k_g1_fft_stage", "unit": "T IMAD.WIDE lane-ops/s (32x32+64)",
                         "achieved": round(imad_meas / 1e12, 3) if imad_meas else None, "peak": round(imad_peak / 1e12, 3),
                         "frac": round(imad_meas / imad_peak, 4) if imad_meas else None,
                         "how": "ncu inst_executed x IMAD.WIDE share x 32 per launch (profiles/stage_kernel_traffic.json) / live launch duration; "
                                "peak = 148 SMs x 4 sub-partitions x 8 lane-ops per cycle x %.0f MHz sampled under load" % sm_mhz,
                         "ncu_fmaheavy_pipe_busy_pct": tj.get("fmaheavy_pipe_busy_pct") if tj else None,
                         "model": {"achieved": round(imad_model / 1e12, 3), "peak": round(fp_peak * IMAD_PER_MUL / 1e12, 3),
                                   "frac": round(imad_model / (fp_peak * IMAD_PER_MUL), 4), "fp_mul_probe_G_per_s": round(fp_peak / 1e9, 3),
                                   "fp_mul_per_blob": round(n_mul), "fp_sqr_per_blob": round(n_sqr),
                                   "how": "whole step: model count of Fp products/squares per blob x (300 | 234) multiply-adds x blobs/s per GPU "
                                          "vs b200_probe_fp_mul (dependent Montgomery products, all SMs) x 300"}},
        "kernels": kernels,
        "kernel_class_ms_per_step": {k: round(cls_ms[i] / K, 3) for i, k in enumerate(PROFILE_CLASSES)},
    }
    if not args.no_components:
        out["components"] = components
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline()
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------- CPU legs
_ORC = {}


def _oracle_fk():
    """Oracle (oracle/kzg_oracle.c: C restatement of the reference algorithm) set up for n = 4096."""
    if "fk" not in _ORC:
        from oracle import cref
        cref.build()
        setup = cref.generate_setup_g1(SECRET, 2 * N_COEFFS)
        _ORC["fk"] = cref.FK20(SCALE, setup, 2 * N_COEFFS)
    return _ORC["fk"]


def _cpu_sample(fk, cores, nblobs):
    """Times the oracle on `nblobs` benchmark blobs: one blob per thread when there are at least `cores` blobs
    (the reference is single threaded, kzg.go / fk20_single.go; independent blobs are the fair way to use every
    core), threads inside the transforms otherwise."""
    from go_kzg_b200.synth import blob_polys
    polys = blob_polys(nblobs, N_COEFFS)
    t0 = time.perf_counter()
    fk.commit_fk20_batch(polys, cores)
    return time.perf_counter() - t0


def cpu_baseline(_polys=None):
    from oracle import cref
    fk = _oracle_fk()
    cores = cref.max_threads()
    dt = _cpu_sample(fk, cores, cores)
    return {"value": round(cores / dt, 5), "unit": "blobs/s", "cores": cores, "kind": "port",
            "sample": "%d blobs of n=%d (one per thread), CommitToPoly + FK20Single, reference algorithm restated in C "
                      "(oracle/kzg_oracle.c); %.1f s" % (cores, N_COEFFS, dt)}


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host cores.  The reference is pure Go on
    un-vendored modules and no Go toolchain exists in the image, so this times the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cref
    fk = _oracle_fk()
    cores = cref.max_threads()
    # calibration (doubles as the warm-up): one blob with the threads inside the transforms
    t1 = _cpu_sample(fk, cores, 1)
    # a step = one blob per thread when the whole run then stays within a few minutes, else a single blob
    est_parallel = t1 * cores
    nb = cores if args.steps * est_parallel <= 180.0 else 1
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _cpu_sample(fk, cores, nb)
    dt = time.perf_counter() - t0
    v = args.steps * nb / dt
    sample = ("%d blob(s) per step (%s), reference algorithm restated in C (oracle/), %d threads"
              % (nb, "one per thread" if nb > 1 else "threads inside the transforms", cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": "blobs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": 1, "ms_per_step": round(dt / args.steps * 1e3, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "CommitToPoly + FK20Single, n=%d, %d blob(s) per step (bounded sample), all host threads" % (N_COEFFS, nb)},
        "cpu_baseline": {"value": round(v, 5), "unit": "blobs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 5), "unit": "blobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="blobs per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-components", action="store_true", help="skip the extra component measurements (configs 4 and 5, one-polynomial latency)")
    ap.add_argument("--config5-scale", type=int, default=21, help="FFT scale of the config 5 component (21: n = 2^20 coefficients); 0 skips it")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
