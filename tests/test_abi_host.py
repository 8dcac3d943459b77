"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/b200_kzg.h
declares, and its host-side level-1 operations (package bls semantics: bls/bignum_kilic.go,
bls/bls_kilic.go) agree with the oracle.  No compute entry point is called (no GPU here)."""
import os
import random
import re

import numpy as np
import pytest

import go_kzg_b200 as kzg
from go_kzg_b200 import build as kbuild
from oracle import cref, pyref

R = pyref.R_MOD
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    kbuild.build()
    return kzg.lib()


def _p(a):
    return a.ctypes.data


def test_exports_every_declared_symbol(L):
    hdr = open(os.path.join(ROOT, "include", "b200_kzg.h")).read()
    names = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cuda_device_is_an_error_not_a_fallback(L):
    if L.b200_device_count() > 0:
        pytest.skip("a GPU is present")
    import ctypes as C
    h = C.c_void_p()
    assert L.b200_fft_settings_new(4, C.byref(h)) == kzg.kzg.NO_DEVICE
    with pytest.raises(kzg.B200Error):
        kzg.FFTSettings(4)
    out = np.zeros(18, dtype=np.uint64)
    pts = np.zeros((2, 18), dtype=np.uint64)
    sc = np.zeros((2, 4), dtype=np.uint64)
    assert L.b200_g1_lincomb(_p(pts), _p(sc), 2, _p(out)) == kzg.kzg.NO_DEVICE


def test_fr_ops_vs_python_ints(L):
    rng = random.Random(1)
    vals = [0, 1, 2, R - 1, R - 2] + [rng.randrange(R) for _ in range(40)]
    for a in vals:
        for b in (vals[0], vals[3], vals[7], vals[11]):
            A, B = kzg.fr_from_ints([a]), kzg.fr_from_ints([b])
            out = np.zeros((1, 4), dtype=np.uint64)
            L.b200_fr_add(_p(out), _p(A), _p(B)); assert kzg.fr_to_ints(out) == [(a + b) % R]
            L.b200_fr_sub(_p(out), _p(A), _p(B)); assert kzg.fr_to_ints(out) == [(a - b) % R]
            L.b200_fr_mul(_p(out), _p(A), _p(B)); assert kzg.fr_to_ints(out) == [(a * b) % R]
            if b:
                L.b200_fr_div(_p(out), _p(A), _p(B)); assert kzg.fr_to_ints(out) == [a * pow(b, -1, R) % R]
        A = kzg.fr_from_ints([a])
        out = np.zeros((1, 4), dtype=np.uint64)
        L.b200_fr_inv(_p(out), _p(A))
        assert kzg.fr_to_ints(out) == [pow(a, -1, R) if a else 0]


def test_fr_aliasing(L):
    """bls/bignum_test.go:8-70: dst may alias either operand."""
    a, b = 1234567891011121314151617181920, R - 5
    A, B = kzg.fr_from_ints([a]), kzg.fr_from_ints([b])
    L.b200_fr_mul(_p(A), _p(A), _p(B)); assert kzg.fr_to_ints(A) == [a * b % R]
    A = kzg.fr_from_ints([a])
    L.b200_fr_add(_p(B), _p(A), _p(B)); assert kzg.fr_to_ints(B) == [(a + b) % R]
    A = kzg.fr_from_ints([a])
    L.b200_fr_mul(_p(A), _p(A), _p(A)); assert kzg.fr_to_ints(A) == [a * a % R]


def test_div_mod_fr(L):
    """bls/bignum_test.go:73-89 TestDivModFr: (a / b) * b == a"""
    a, b = 222222222, 1111
    A, B = kzg.fr_from_ints([a]), kzg.fr_from_ints([b])
    q = np.zeros((1, 4), dtype=np.uint64)
    L.b200_fr_div(_p(q), _p(A), _p(B))
    back = np.zeros((1, 4), dtype=np.uint64)
    L.b200_fr_mul(_p(back), _p(q), _p(B))
    assert kzg.fr_to_ints(back) == [a]


def test_batch_inv(L):
    rng = random.Random(5)
    vals = [rng.randrange(R) for _ in range(33)]
    vals[4] = 0
    a = kzg.fr_from_ints(vals)
    L.b200_fr_batch_inv(_p(a), len(vals))
    assert kzg.fr_to_ints(a) == [pow(v, -1, R) if v else 0 for v in vals]


def test_valid_fr(L):
    """bls/bignum_test.go:91-116 TestValidFr"""
    def ok(v):
        return L.b200_fr_valid(_p(np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8).copy()))
    assert ok(0) and ok(1) and ok(R - 1)
    assert not ok(R) and not ok(R + 1) and not ok((1 << 256) - 1)


def test_roots_of_unity(L, goldens):
    """bls/globals.go:27-60"""
    for k, v in enumerate(goldens["scale2_root_of_unity"]["values"]):
        out = np.zeros((1, 4), dtype=np.uint64)
        L.b200_fr_root_of_unity(k, _p(out))
        assert kzg.fr_to_ints(out) == [int(v)]


def test_g1_host_ops_vs_oracle(L, goldens):
    rng = random.Random(9)
    G = np.zeros(18, dtype=np.uint64)
    L.b200_g1_generator(_p(G))
    assert np.array_equal(G, cref.g1_generator())
    g = goldens["point_compression"]                       # bls/bls_test.go:11-23
    P = np.zeros(18, dtype=np.uint64)
    L.b200_g1_mul(_p(P), _p(G), _p(kzg.fr_from_ints([int(g["scalar"])])))
    assert bytes(kzg.g1_to_compressed(P)[0]) == bytes(g["expected"])
    inf = np.zeros(18, dtype=np.uint64)
    assert bytes(kzg.g1_to_compressed(inf)[0]) == bytes([0xC0]) + bytes(47)
    for _ in range(4):
        a, b = rng.randrange(R), rng.randrange(R)
        A, B, S, D = (np.zeros(18, dtype=np.uint64) for _ in range(4))
        L.b200_g1_mul(_p(A), _p(G), _p(kzg.fr_from_ints([a])))
        L.b200_g1_mul(_p(B), _p(G), _p(kzg.fr_from_ints([b])))
        L.b200_g1_add(_p(S), _p(A), _p(B))
        L.b200_g1_sub(_p(D), _p(A), _p(B))
        want = cref.g1_compress(cref.g1_mul_gen([a, b, a + b, a - b]))
        got = kzg.g1_to_compressed(np.stack([A, B, S, D]))
        assert np.array_equal(got, want)
        assert L.b200_g1_equal(_p(S), _p(cref.g1_mul_gen([a + b])[0].copy())) == 1
        assert L.b200_g1_equal(_p(S), _p(D)) == 0
        # P + P (doubling branch), P - P (infinity), P + inf
        L.b200_g1_add(_p(S), _p(A), _p(A))
        assert np.array_equal(kzg.g1_to_compressed(S), cref.g1_compress(cref.g1_mul_gen([2 * a])))
        L.b200_g1_sub(_p(S), _p(A), _p(A))
        assert L.b200_g1_equal(_p(S), _p(inf)) == 1
        L.b200_g1_add(_p(S), _p(A), _p(inf))
        assert L.b200_g1_equal(_p(S), _p(A)) == 1
        N = A.copy()
        L.b200_g1_neg(_p(N))
        assert np.array_equal(kzg.g1_to_compressed(N), cref.g1_compress(cref.g1_mul_gen([R - a])))


def test_compression_round_trip_on_fixture(L, trusted_setup_bytes):
    """eth/trusted_setup.json setup_G1: decompress -> compress is the identity; values match the oracle."""
    s1, _ = trusted_setup_bytes
    dec = kzg.g1_from_compressed(s1[:64])
    assert np.array_equal(kzg.g1_to_compressed(dec), s1[:64])
    assert np.array_equal(dec, cref.g1_decompress(s1[:64]))
    bad = s1[0].copy()
    bad[0] &= 0x7F                                           # compression flag cleared
    with pytest.raises(kzg.KZGError):
        kzg.g1_from_compressed(bad)


def test_synthetic_generator_is_pinned():
    from kzg_test_util import random_fr_ints
    from go_kzg_b200.synth import random_fr_limbs
    assert kzg.fr_to_ints(random_fr_limbs(300, 0xB2000000)) == random_fr_ints(300, 0xB2000000)
    assert random_fr_ints(2, 0xB2000000) == [
        int(x) for x in kzg.fr_to_ints(random_fr_limbs(2, 0xB2000000))]


def test_device_scalar_multiplication_code_on_the_host(tmp_path):
    """g1_mul_digits (the effective-affine table routine the G1 FFT kernels run) is __host__
    __device__: tools/host_check.cu runs it on the CPU against plain double-and-add for both digit
    recodings, Jacobian inputs with Z != 1 and edge scalars (1, 2, 17, r - 1)."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_check")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "go_kzg_b200", "csrc"), "-I",
                           os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "host_check.cu")],
                          stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout


def test_fp64_pipe_product_matches_integer_product_on_the_host(tmp_path):
    """field_fp64_impl.cuh (Montgomery product in 24-bit limbs on doubles) against fe_mul / fe_sqr:
    random operands and the edge values 0, 1, p - 1, p - 2, 2^380 - 1 (tools/fp64_check.cu)."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fp64_check")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "go_kzg_b200", "csrc"), "-I",
                           os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "fp64_check.cu")],
                          stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout


def test_karatsuba_product_matches_integer_product_on_the_host(tmp_path):
    """field_karatsuba.cuh (Karatsuba product half + m * p-only reduction rows, opt-in) against fe_mul / the
    portable product: random operands and edge values incl. equal halves (tools/kara_check.cu)."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    for levels in ("1", "2"):
        exe = str(tmp_path / ("kara_check" + levels))
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-DB200_KARA_LEVELS=" + levels, "-I", os.path.join(ROOT, "go_kzg_b200", "csrc"),
                               "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "kara_check.cu")],
                              stderr=subprocess.DEVNULL)
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout


def test_from_compressed_rejects_points_outside_the_subgroup(L):
    """kilic's FromCompressed (behind bls/bls_kilic.go:118-121) rejects encodings of curve points that are not in
    the prime-order subgroup.  The library tests membership with the endomorphism ([z^2] P == (beta x, -y)), the
    oracle with r P == infinity: two independent routes that must agree on every candidate x."""
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    found_bad = 0
    for x in range(1, 40):
        rhs = (x * x * x + 4) % P_MOD
        if pow(rhs, (P_MOD - 1) // 2, P_MOD) != 1:
            continue                                     # not on the curve: both reject (checked below)
        enc = np.frombuffer(bytes([0x80]) + x.to_bytes(48, "big")[1:], dtype=np.uint8).copy()
        out = np.zeros(18, dtype=np.uint64)
        rc = L.b200_g1_from_compressed(_p(out), _p(enc))
        try:
            cref.g1_decompress(enc[None, :])
            oracle_ok = True
        except ValueError:
            oracle_ok = False
        assert (rc == 0) == oracle_ok
        found_bad += rc != 0
    assert found_bad >= 5                                # the cofactor is ~2^126: small x never land in G1
    # members are still accepted: the generator and a few multiples round-trip
    pts = cref.g1_mul_gen([1, 2, 1337, R - 1])
    enc = cref.g1_compress(pts)
    assert np.array_equal(kzg.g1_to_compressed(kzg.g1_from_compressed(enc)), enc)
    off = np.frombuffer(bytes([0x80]) + (3).to_bytes(48, "big")[1:], dtype=np.uint8).copy()
    if L.b200_g1_from_compressed(_p(np.zeros(18, dtype=np.uint64)), _p(off)) != 0:
        with pytest.raises(kzg.KZGError):
            kzg.g1_from_compressed(off[None, :])


def test_bucket_msm_pipeline_replayed_on_the_host(tmp_path):
    """The per-thread bodies of the Pippenger pipeline (msm.cuh: GLV split, signed window digits, counting sort,
    bucket accumulation, segment reduction) are __host__ __device__: tools/msm_host_check.cu replays the kernels one
    thread at a time on the CPU against double-and-add -- affine and Jacobian inputs, zero / one / r - 1 scalars, a
    point at infinity, the same term twice (doubling inside a bucket) and P with -P (cancellation)."""
    import shutil
    import subprocess
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "msm_host_check")
    subprocess.check_call([nvcc, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "go_kzg_b200", "csrc"), "-I",
                           os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tools", "msm_host_check.cu")],
                          stderr=subprocess.DEVNULL)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout
    assert "plan n=4096: c=10 W=13 B=512" in out.stdout
