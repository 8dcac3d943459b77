"""ctypes front-end of the C oracle (oracle/kzg_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Encodings are those of include/b200_kzg.h:
Fr = 4 x u64 LE canonical limbs; G1 = Jacobian 18 x u64 canonical limbs (Z == 0 <=> infinity).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkzg_oracle.so")

R_MOD = 52435875175126190479447740508185965837690552500527637822603658699938581184513


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (no GPU needed)."""
    src = os.path.join(_HERE, "kzg_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int, C.c_uint
        L.orc_fs_new.restype = vp
        L.orc_fs_new.argtypes = [u32]
        L.orc_fs_free.argtypes = [vp]
        L.orc_fft_fr.argtypes = [vp, vp, u64, i32, vp]
        L.orc_fft_g1.argtypes = [vp, vp, u64, i32, vp]
        L.orc_g1_generator.argtypes = [vp]
        L.orc_g1_mul.argtypes = [vp, vp, vp]
        L.orc_g1_add.argtypes = [vp, vp, vp]
        L.orc_g1_sub.argtypes = [vp, vp, vp]
        L.orc_g1_equal.argtypes = [vp, vp]
        L.orc_g1_compress.argtypes = [vp, vp]
        L.orc_g1_compress_many.argtypes = [vp, vp, u64]
        L.orc_g1_decompress.argtypes = [vp, vp]
        L.orc_g1_decompress_many.argtypes = [vp, vp, u64]
        L.orc_g1_mul_gen_many.argtypes = [vp, vp, u64]
        L.orc_generate_setup_g1.argtypes = [vp, u64, vp]
        L.orc_lincomb_g1.argtypes = [vp, vp, u64, vp]
        L.orc_fk20_new.restype = vp
        L.orc_fk20_new.argtypes = [u32, vp, u64, u64, u64]
        L.orc_fk20_free.argtypes = [vp]
        L.orc_fk20_x_ext_fft.argtypes = [vp, u64, vp]
        L.orc_commit_to_poly.argtypes = [vp, vp, u64, vp]
        L.orc_fk20_single.argtypes = [vp, vp, u64, i32, vp]
        L.orc_fk20_multi_da.argtypes = [vp, vp, u64, vp]
        L.orc_commit_fk20_batch.argtypes = [vp, vp, u64, u64, vp, vp, i32]
        L.orc_das_fft_extension.argtypes = [vp, vp, u64]
        L.orc_zero_poly.argtypes = [vp, vp, u64, u64, vp, vp]
        L.orc_recover_poly_from_samples.argtypes = [vp, vp, vp, u64, vp]
        L.orc_set_threads.argtypes = [i32]
        L.orc_init()
        _lib = L
    return _lib


# ---------------------------------------------------------------- conversions
def fr_to_limbs(vals) -> np.ndarray:
    """list of ints (already reduced mod r) -> (n,4) uint64"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = int(v)
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_fr(a: np.ndarray) -> list:
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]


def affine_to_g1(points) -> np.ndarray:
    """list of (x, y) int pairs or None -> (n,18) uint64 Jacobian with Z=1"""
    out = np.zeros((len(points), 18), dtype=np.uint64)
    for i, p in enumerate(points):
        if p is None:
            continue
        for c, v in enumerate((p[0], p[1], 1)):
            for j in range(6):
                out[i, 6 * c + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype=np.uint64):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------- wrappers
class FFTSettings:
    """fft.go:34-61"""

    def __init__(self, max_scale: int):
        self.max_scale = max_scale
        self.max_width = 1 << max_scale
        self.h = lib().orc_fs_new(max_scale)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fs_free(self.h)
            self.h = None

    def fft(self, vals: np.ndarray, inv: bool = False) -> np.ndarray:
        vals = _c(vals).reshape(-1, 4)
        n = vals.shape[0]
        npow = 1 if n == 0 else 1 << (n - 1).bit_length()
        out = np.zeros((npow, 4), dtype=np.uint64)
        rc = lib().orc_fft_fr(self.h, _p(vals), n, int(inv), _p(out))
        if rc:
            raise ValueError("orc_fft_fr rc=%d" % rc)
        return out

    def fft_g1(self, pts: np.ndarray, inv: bool = False) -> np.ndarray:
        pts = _c(pts).reshape(-1, 18)
        out = np.zeros_like(pts)
        rc = lib().orc_fft_g1(self.h, _p(pts), pts.shape[0], int(inv), _p(out))
        if rc:
            raise ValueError("orc_fft_g1 rc=%d" % rc)
        return out

    def das_fft_extension(self, vals: np.ndarray) -> np.ndarray:
        v = _c(vals).reshape(-1, 4).copy()
        rc = lib().orc_das_fft_extension(self.h, _p(v), v.shape[0])
        if rc:
            raise RuntimeError("orc_das_fft_extension rc=%d" % rc)
        return v

    def zero_poly(self, missing, length):
        m = _c(np.asarray(missing, dtype=np.uint64))
        ze = np.zeros((length, 4), dtype=np.uint64)
        zp = np.zeros((length, 4), dtype=np.uint64)
        rc = lib().orc_zero_poly(self.h, _p(m), len(missing), length, _p(ze), _p(zp))
        if rc:
            raise RuntimeError("orc_zero_poly rc=%d" % rc)
        return ze, zp

    def recover(self, samples: np.ndarray, present: np.ndarray) -> np.ndarray:
        s = _c(samples).reshape(-1, 4)
        pr = _c(present, np.uint8)
        out = np.zeros_like(s)
        rc = lib().orc_recover_poly_from_samples(self.h, _p(s), _p(pr), s.shape[0], _p(out))
        if rc:
            raise RuntimeError("orc_recover rc=%d" % rc)
        return out


def g1_generator() -> np.ndarray:
    out = np.zeros(18, dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_mul(p: np.ndarray, k: int) -> np.ndarray:
    out = np.zeros(18, dtype=np.uint64)
    kk = fr_to_limbs([k % R_MOD])
    lib().orc_g1_mul(_p(out), _p(_c(p)), _p(kk))
    return out


def g1_add(a, b) -> np.ndarray:
    out = np.zeros(18, dtype=np.uint64)
    lib().orc_g1_add(_p(out), _p(_c(a)), _p(_c(b)))
    return out


def g1_sub(a, b) -> np.ndarray:
    out = np.zeros(18, dtype=np.uint64)
    lib().orc_g1_sub(_p(out), _p(_c(a)), _p(_c(b)))
    return out


def g1_equal(a, b) -> bool:
    return bool(lib().orc_g1_equal(_p(_c(a)), _p(_c(b))))


def g1_compress(pts: np.ndarray) -> np.ndarray:
    """(n,18) -> (n,48) uint8"""
    pts = _c(pts).reshape(-1, 18)
    out = np.zeros((pts.shape[0], 48), dtype=np.uint8)
    lib().orc_g1_compress_many(_p(out), _p(pts), pts.shape[0])
    return out


def g1_decompress(b: np.ndarray) -> np.ndarray:
    b = _c(b, np.uint8).reshape(-1, 48)
    out = np.zeros((b.shape[0], 18), dtype=np.uint64)
    if lib().orc_g1_decompress_many(_p(out), _p(b), b.shape[0]):
        raise ValueError("bad compressed point")
    return out


def g1_mul_gen(scalars) -> np.ndarray:
    """[k_i * G] for ints k_i (the exponent-domain oracle's last step)."""
    s = fr_to_limbs([int(k) % R_MOD for k in scalars])
    out = np.zeros((len(scalars), 18), dtype=np.uint64)
    lib().orc_g1_mul_gen_many(_p(out), _p(s), len(scalars))
    return out


def generate_setup_g1(secret: int, n: int) -> np.ndarray:
    """setup.go:9-26"""
    out = np.zeros((n, 18), dtype=np.uint64)
    lib().orc_generate_setup_g1(_p(fr_to_limbs([secret % R_MOD])), n, _p(out))
    return out


def lincomb_g1(pts: np.ndarray, scalars: np.ndarray) -> np.ndarray:
    pts = _c(pts).reshape(-1, 18)
    sc = _c(scalars).reshape(-1, 4)
    assert pts.shape[0] == sc.shape[0]
    out = np.zeros(18, dtype=np.uint64)
    lib().orc_lincomb_g1(_p(pts), _p(sc), pts.shape[0], _p(out))
    return out


class FK20:
    """KZGSettings + FK20SingleSettings / FK20MultiSettings (kzg.go:11-116)."""

    def __init__(self, max_scale: int, secret_g1: np.ndarray, n2: int, chunk_len: int = 1):
        sg = _c(secret_g1).reshape(-1, 18)
        self.n2, self.chunk_len = n2, chunk_len
        self.h = lib().orc_fk20_new(max_scale, _p(sg), sg.shape[0], n2, chunk_len)
        if not self.h:
            raise RuntimeError("orc_fk20_new: the reference would panic on these arguments")

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fk20_free(self.h)
            self.h = None

    def x_ext_fft(self, file: int = 0) -> np.ndarray:
        out = np.zeros((self.n2 // self.chunk_len, 18), dtype=np.uint64)
        lib().orc_fk20_x_ext_fft(self.h, file, _p(out))
        return out

    def commit(self, coeffs: np.ndarray) -> np.ndarray:
        c = _c(coeffs).reshape(-1, 4)
        out = np.zeros(18, dtype=np.uint64)
        lib().orc_commit_to_poly(self.h, _p(c), c.shape[0], _p(out))
        return out

    def fk20_single(self, poly: np.ndarray, da: bool = False) -> np.ndarray:
        p = _c(poly).reshape(-1, 4)
        n = p.shape[0]
        out = np.zeros(((2 * n) if da else n, 18), dtype=np.uint64)
        rc = lib().orc_fk20_single(self.h, _p(p), n, int(da), _p(out))
        if rc:
            raise RuntimeError("orc_fk20_single rc=%d" % rc)
        return out

    def fk20_multi_da(self, poly: np.ndarray) -> np.ndarray:
        p = _c(poly).reshape(-1, 4)
        n = p.shape[0]
        out = np.zeros((2 * n // self.chunk_len, 18), dtype=np.uint64)
        rc = lib().orc_fk20_multi_da(self.h, _p(p), n, _p(out))
        if rc:
            raise RuntimeError("orc_fk20_multi_da rc=%d" % rc)
        return out

    def commit_fk20_batch(self, polys: np.ndarray, nthreads: int = 0):
        p = _c(polys)
        nb, n = p.shape[0], p.shape[1]
        commits = np.zeros((nb, 18), dtype=np.uint64)
        proofs = np.zeros((nb, n, 18), dtype=np.uint64)
        rc = lib().orc_commit_fk20_batch(self.h, _p(p), n, nb, _p(commits), _p(proofs), nthreads)
        if rc:
            raise RuntimeError("orc_commit_fk20_batch rc=%d" % rc)
        return commits, proofs


def max_threads() -> int:
    return lib().orc_max_threads()
