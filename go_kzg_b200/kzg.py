"""ctypes mirror of the reference's Go API over the C ABI of libb200kzg.so.

Encodings (include/b200_kzg.h): Fr = (n, 4) uint64 little-endian canonical limbs;
G1 = (n, 18) uint64 Jacobian canonical limbs, Z == 0 <=> infinity; compressed G1 = (n, 48) uint8.

Error behaviour follows the reference: functions that return `error` in Go raise KZGError,
functions that `panic` in Go raise KZGPanic; CUDA problems raise B200Error.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

R_MOD = 52435875175126190479447740508185965837690552500527637822603658699938581184513

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("B200_KZG_LIB") or os.path.join(_HERE, "lib", "libb200kzg.so")   # override: tuning builds only

OK, TOO_LARGE, NOT_POW2, LEN_MISMATCH, BAD_INPUT, ERR_CUDA, NO_DEVICE, TOO_SMALL, RECOVERY, ZERO_EVAL = range(10)


class B200Error(RuntimeError):
    """CUDA / device failure inside the library (no CPU fallback exists)."""


class KZGError(ValueError):
    """The reference returns an `error` for this condition."""

    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


class KZGPanic(Exception):
    """The reference panics on this condition."""

    def __init__(self, status, msg):
        super().__init__(msg)
        self.status = status


def lib_path() -> str:
    return _LIB_PATH


_lib = None


def lib():
    """Loads libb200kzg.so (built by go_kzg_b200/build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise B200Error("libb200kzg.so is not built: run `python -m go_kzg_b200.build` "
                        "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(_LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.b200_strerror.restype = C.c_char_p
    L.b200_strerror.argtypes = [i32]
    L.b200_last_cuda_error.restype = C.c_char_p
    L.b200_device_count.restype = i32
    L.b200_set_device.argtypes = [i32]
    for name in ("b200_fr_add", "b200_fr_sub", "b200_fr_mul", "b200_fr_div", "b200_g1_add", "b200_g1_sub", "b200_g1_mul"):
        getattr(L, name).argtypes = [vp, vp, vp]
        getattr(L, name).restype = None
    L.b200_fr_inv.argtypes = [vp, vp]
    L.b200_fr_inv.restype = None
    L.b200_fr_batch_inv.argtypes = [vp, sz]
    L.b200_fr_batch_inv.restype = None
    L.b200_fr_valid.argtypes = [vp]
    L.b200_fr_root_of_unity.argtypes = [C.c_uint, vp]
    L.b200_fr_root_of_unity.restype = None
    L.b200_g1_generator.argtypes = [vp]
    L.b200_g1_generator.restype = None
    L.b200_g1_neg.argtypes = [vp]
    L.b200_g1_neg.restype = None
    L.b200_g1_equal.argtypes = [vp, vp]
    L.b200_g1_to_compressed.argtypes = [vp, vp]
    L.b200_g1_to_compressed.restype = None
    L.b200_g1_from_compressed.argtypes = [vp, vp]
    L.b200_g1_to_compressed_many.argtypes = [vp, vp, sz]
    L.b200_g1_to_compressed_many.restype = None
    L.b200_g1_from_compressed_many.argtypes = [vp, vp, sz]
    L.b200_g1_from_compressed_batch.argtypes = [vp, sz, vp, vp]
    L.b200_toeplitz_part2.argtypes = [vp, vp, sz, vp, sz, vp]
    L.b200_toeplitz_part3.argtypes = [vp, vp, sz, vp]
    L.b200_evaluate_poly_in_evaluation_form_batch.argtypes = [vp, vp, vp, sz, sz, i32, vp]
    L.b200_check_proof_single_g1_batch.argtypes = [vp, vp, sz, vp]
    L.b200_check_proof_multi_g1_batch.argtypes = [vp, vp, vp, vp, sz, sz, vp, vp]
    for name in ("b200_g2_add", "b200_g2_sub", "b200_g2_mul", "b200_g2_to_compressed"):
        getattr(L, name).argtypes = [vp, vp, vp] if name != "b200_g2_to_compressed" else [vp, vp]
        getattr(L, name).restype = None
    L.b200_g1_to_affine.argtypes = [vp, vp]
    L.b200_g1_to_affine.restype = None
    L.b200_g2_to_affine.argtypes = [vp, vp]
    L.b200_g2_to_affine.restype = None
    L.b200_g2_generator.argtypes = [vp]
    L.b200_g2_generator.restype = None
    L.b200_g2_neg.argtypes = [vp]
    L.b200_g2_neg.restype = None
    L.b200_g2_equal.argtypes = [vp, vp]
    L.b200_g2_from_compressed.argtypes = [vp, vp]
    L.b200_generate_testing_setup_g2.argtypes = [vp, sz, vp]
    L.b200_pairings_verify.argtypes = [vp, vp, vp, vp, C.POINTER(i32)]
    L.b200_pairing.argtypes = [vp, vp, vp]
    L.b200_kzg_settings_set_secret_g2.argtypes = [vp, vp, sz]
    L.b200_check_proof_single.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i32)]
    L.b200_check_proof_single_batch.argtypes = [vp, vp, vp, vp, vp, sz, vp]
    L.b200_check_proof_multi.argtypes = [vp, vp, vp, vp, vp, sz, C.POINTER(i32)]
    L.b200_check_proof_multi_batch.argtypes = [vp, vp, vp, vp, vp, sz, sz, vp]
    L.b200_check_proof_single_aggregate.argtypes = [vp, vp, vp, vp, vp, vp, sz, C.POINTER(i32)]
    L.b200_check_proof_multi_aggregate.argtypes = [vp, vp, vp, vp, vp, sz, vp, sz, C.POINTER(i32)]
    L.b200_g1_lincomb.argtypes = [vp, vp, sz, vp]
    L.b200_g1_mul_many.argtypes = [vp, vp, sz, vp]
    L.b200_fft_settings_new.argtypes = [C.c_uint8, C.POINTER(vp)]
    L.b200_fft_settings_free.argtypes = [vp]
    L.b200_fft_settings_free.restype = None
    L.b200_fs_max_width.argtypes = [vp]
    L.b200_fs_max_width.restype = u64
    L.b200_fs_roots.argtypes = [vp, i32, vp]
    L.b200_fft_fr.argtypes = [vp, vp, sz, i32, vp]
    L.b200_fft_fr_batch.argtypes = [vp, vp, sz, sz, i32, vp]
    L.b200_inplace_fft_fr.argtypes = [vp, vp, sz, i32, vp]
    L.b200_fft_g1.argtypes = [vp, vp, sz, i32, vp]
    L.b200_fft_g1_batch.argtypes = [vp, vp, sz, sz, i32, vp]
    L.b200_das_fft_extension.argtypes = [vp, vp, sz]
    L.b200_das_fft_extension_batch.argtypes = [vp, vp, sz, sz]
    L.b200_das_fft_extension_g1.argtypes = [vp, vp, sz]
    L.b200_zero_poly_via_multiplication.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_recover_poly_from_samples.argtypes = [vp, vp, vp, sz, vp]
    L.b200_recover_poly_from_samples_batch.argtypes = [vp, vp, vp, sz, sz, vp]
    L.b200_kzg_settings_new.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
    L.b200_kzg_settings_free.argtypes = [vp]
    L.b200_kzg_settings_free.restype = None
    L.b200_commit_to_poly.argtypes = [vp, vp, sz, vp]
    L.b200_commit_to_poly_batch.argtypes = [vp, vp, sz, sz, vp]
    L.b200_fk20_single_settings_new.argtypes = [vp, sz, C.POINTER(vp)]
    L.b200_fk20_multi_settings_new.argtypes = [vp, sz, sz, C.POINTER(vp)]
    L.b200_fk20_multi_settings_new_sharded.argtypes = [vp, sz, sz, sz, sz, C.POINTER(vp)]
    L.b200_fk20_settings_new_from_x_ext_fft.argtypes = [vp, sz, sz, sz, sz, vp, C.POINTER(vp)]
    L.b200_fk20_settings_free.argtypes = [vp]
    L.b200_fk20_settings_free.restype = None
    L.b200_fk20_x_ext_fft.argtypes = [vp, sz, vp]
    L.b200_fk20_single.argtypes = [vp, vp, sz, vp]
    L.b200_fk20_single_da_optimized.argtypes = [vp, vp, sz, vp]
    L.b200_da_using_fk20.argtypes = [vp, vp, sz, vp]
    L.b200_da_using_fk20_batch.argtypes = [vp, vp, sz, sz, vp]
    L.b200_da_using_fk20_batch_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_fk20_multi_da_optimized.argtypes = [vp, vp, sz, vp]
    L.b200_da_using_fk20_multi.argtypes = [vp, vp, sz, vp]
    L.b200_commit_fk20_batch.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_commit_fk20_batch_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.b200_commit_fk20_batch_compressed.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_commit_fk20_batch_compressed_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.b200_g1_compress_dev.argtypes = [vp, sz, vp, vp]
    L.b200_g1_to_compressed_batch.argtypes = [vp, sz, vp]
    L.b200_blob_to_kzg_commitment_batch.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_compute_kzg_proof_batch.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp]
    L.b200_fk20_multi_partial_dev.argtypes = [vp, vp, sz, sz, sz, vp, vp]
    L.b200_g1_sum_dev.argtypes = [vp, sz, sz, vp, vp]
    L.b200_fk20_multi_finish_dev.argtypes = [vp, vp, i32, vp, vp]
    L.b200_fk20_multi_finish_local_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_fk20_multi_finish_merge_dev.argtypes = [vp, vp, sz, i32, vp, vp]
    L.b200_fk20_multi_finish_merge_part_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_fk20_multi_finish_assemble_dev.argtypes = [vp, vp, sz, i32, vp, vp]
    L.b200_commit_partial_dev.argtypes = [vp, vp, sz, sz, vp, vp]
    L.b200_generate_testing_setup_g1.argtypes = [vp, sz, vp]
    L.b200_fk20_last_launch_count.argtypes = [vp]
    L.b200_fk20_last_launch_count.restype = u64
    L.b200_last_launch_count.restype = u64
    L.b200_set_fixed_base_window.argtypes = [i32]
    L.b200_set_latency_mode.argtypes = [i32]
    L.b200_selftest_field.argtypes = [sz, u64, C.POINTER(u64)]
    L.b200_probe_fp_mul.argtypes = [sz, i32, C.POINTER(C.c_float)]
    L.b200_profile_end.argtypes = [C.POINTER(C.c_double), C.POINTER(u64)]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fr(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)


def _g1(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 18)


def _g2(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 36)


def _raise(status: int, errors=(), what: str = ""):
    """status -> exception, following the reference's error/panic split for this call."""
    if status == OK:
        return
    msg = "%s: %s" % (what, lib().b200_strerror(status).decode())
    if status in (ERR_CUDA, NO_DEVICE):
        raise B200Error(msg + " [" + lib().b200_last_cuda_error().decode() + "]")
    if status in errors:
        raise KZGError(status, msg)
    raise KZGPanic(status, msg)


# ------------------------------------------------------------------------------- conversions
def fr_from_ints(vals) -> np.ndarray:
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        v = int(v)
        for j in range(4):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def fr_to_ints(a) -> list:
    a = _fr(a)
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]


def g1_to_compressed(pts) -> np.ndarray:
    """bls/bls_kilic.go:114 ToCompressedG1 over an array (host side)."""
    pts = _g1(pts)
    out = np.zeros((pts.shape[0], 48), dtype=np.uint8)
    lib().b200_g1_to_compressed_many(_p(out), _p(pts), pts.shape[0])
    return out


def g1_to_compressed_device(pts) -> np.ndarray:
    """ToCompressedG1 over an array, normalised and compressed on the GPU (b200_g1_to_compressed_batch)."""
    pts = _g1(pts)
    out = np.zeros((pts.shape[0], 48), dtype=np.uint8)
    _raise(lib().b200_g1_to_compressed_batch(_p(pts), pts.shape[0], _p(out)), what="ToCompressedG1 (device)")
    return out


def g1_from_compressed(b) -> np.ndarray:
    """bls/bls_kilic.go:118 FromCompressedG1 over an array (host side); error on bad encodings."""
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 48)
    out = np.zeros((b.shape[0], 18), dtype=np.uint64)
    _raise(lib().b200_g1_from_compressed_many(_p(out), _p(b), b.shape[0]), errors=(BAD_INPUT,), what="FromCompressedG1")
    return out


def g1_from_compressed_device(b, return_ok: bool = False):
    """FromCompressedG1 over an array on the GPU (b200_g1_from_compressed_batch): square root and subgroup check per
    point on the device.  Raises KZGError on a bad encoding unless return_ok, which hands back (points, ok)."""
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 48)
    out = np.zeros((b.shape[0], 18), dtype=np.uint64)
    ok = np.zeros(b.shape[0], dtype=np.uint8)
    rc = lib().b200_g1_from_compressed_batch(_p(b), b.shape[0], _p(out), _p(ok))
    if return_ok and rc in (OK, BAD_INPUT):
        return out, ok.astype(bool)
    _raise(rc, errors=(BAD_INPUT,), what="FromCompressedG1 (device)")
    return out


def g1_marshal_text(pts) -> list:
    """bls/bls_all.go:20-23 (*G1Point).MarshalText: hex of the compressed encoding, no 0x prefix."""
    return [bytes(row).hex() for row in g1_to_compressed(pts)]


def g1_unmarshal_text(texts, device: bool = True) -> np.ndarray:
    """bls/bls_all.go:25-39 (*G1Point).UnmarshalText over a list: hex (no 0x prefix) -> FromCompressedG1; error on bad
    hex, a wrong length or a bad point."""
    raw = []
    for t in texts:
        if isinstance(t, bytes):
            t = t.decode()
        try:
            v = bytes.fromhex(t)
        except ValueError as e:
            raise KZGError(BAD_INPUT, "UnmarshalText: %s" % e)
        if len(v) != 48:
            raise KZGError(BAD_INPUT, "UnmarshalText: %d bytes, want 48" % len(v))
        raw.append(v)
    arr = np.frombuffer(b"".join(raw), dtype=np.uint8).reshape(-1, 48) if raw else np.zeros((0, 48), dtype=np.uint8)
    return g1_from_compressed_device(arr) if device else g1_from_compressed(arr)


def bit_reversal_permutation(a: np.ndarray) -> np.ndarray:
    """eth/helpers.go bitReversalPermutation on the first axis (length a power of two)."""
    n = a.shape[0]
    bits = n.bit_length() - 1
    idx = np.arange(n, dtype=np.uint64)
    rev = np.zeros(n, dtype=np.uint64)
    for k in range(bits):
        rev |= ((idx >> np.uint64(k)) & np.uint64(1)) << np.uint64(bits - 1 - k)
    return a[rev.astype(np.int64)]


def load_trusted_setup(source, device: bool = True) -> dict:
    """eth/globals.go:33-49: parses a trusted_setup.json-shaped document ({"setup_G1": [...], "setup_G2": [...],
    "setup_G1_lagrange": [...]}, hex strings without 0x).  Returns {"setup_G1": (n, 18) points,
    "setup_G1_lagrange": (n, 18) points ALREADY in reverse bit order (eth/globals.go:48 kzgSetupLagrange),
    "setup_G2": (m, 96) raw compressed bytes (g2_from_compressed decodes the entries a verifier needs; a decode includes
    the subgroup check, ~ms per point on the host)}.
    `source` is a path, a JSON string / bytes, or an already parsed dict."""
    import json
    if isinstance(source, dict):
        doc = source
    else:
        if isinstance(source, bytes):
            source = source.decode()
        if isinstance(source, str) and not source.lstrip().startswith("{"):
            with open(source) as f:
                source = f.read()
        doc = json.loads(source)
    g1 = g1_unmarshal_text(doc.get("setup_G1", []), device)
    lag = g1_unmarshal_text(doc.get("setup_G1_lagrange", []), device)
    g2 = doc.get("setup_G2", [])
    for t in g2:
        if len(bytes.fromhex(t)) != 96:
            raise KZGError(BAD_INPUT, "setup_G2: want 96-byte compressed points")
    g2b = np.frombuffer(b"".join(bytes.fromhex(t) for t in g2), dtype=np.uint8).reshape(-1, 96) if g2 else np.zeros((0, 96), dtype=np.uint8)
    return {"setup_G1": g1, "setup_G1_lagrange": bit_reversal_permutation(lag) if lag.shape[0] else lag, "setup_G2": g2b}


def check_proof_single_g1(commitments, ys) -> np.ndarray:
    """kzg_single_proofs.go:57-75 CheckProofSingle, G1 side for a batch: commitment - y G.  The pairing
    e(result, [1]_2) == e(proof, [s - x]_2) stays with the caller's backend."""
    c, y = _g1(commitments), _fr(ys)
    assert c.shape[0] == y.shape[0]
    out = np.zeros_like(c)
    _raise(lib().b200_check_proof_single_g1_batch(_p(c), _p(y), c.shape[0], _p(out)), what="CheckProofSingle (G1 side)")
    return out


# ---- G2 and the pairing: host code of the library (verification side; bls/bls_kilic.go:69-104,123-130,152-158) ----
def g2_generator() -> np.ndarray:
    out = np.zeros(36, dtype=np.uint64)
    lib().b200_g2_generator(_p(out))
    return out


def _g2_binop(name, a, b) -> np.ndarray:
    a, b = _g2(a)[0].copy(), np.ascontiguousarray(b, dtype=np.uint64).reshape(-1)
    out = np.zeros(36, dtype=np.uint64)
    getattr(lib(), name)(_p(out), _p(a), _p(b))
    return out


def g2_add(a, b) -> np.ndarray:
    return _g2_binop("b200_g2_add", a, b)


def g2_sub(a, b) -> np.ndarray:
    return _g2_binop("b200_g2_sub", a, b)


def g2_mul(a, k: int) -> np.ndarray:
    """bls/bls_kilic.go:80-84 MulG2 (k an integer, reduced mod r)"""
    return _g2_binop("b200_g2_mul", a, fr_from_ints([k % R_MOD])[0])


def g2_neg(a) -> np.ndarray:
    out = _g2(a)[0].copy()
    lib().b200_g2_neg(_p(out))
    return out


def g2_equal(a, b) -> bool:
    a, b = _g2(a)[0].copy(), _g2(b)[0].copy()
    return bool(lib().b200_g2_equal(_p(a), _p(b)))


def g2_to_compressed(pts) -> np.ndarray:
    """bls/bls_kilic.go:123-125 ToCompressedG2 per point -> (n, 96) uint8"""
    g = _g2(pts)
    out = np.zeros((g.shape[0], 96), dtype=np.uint8)
    for i in range(g.shape[0]):
        row = np.ascontiguousarray(g[i])
        lib().b200_g2_to_compressed(_p(out[i]), _p(row))
    return out


def g2_from_compressed(b) -> np.ndarray:
    """bls/bls_kilic.go:127-130 FromCompressedG2 per point; a bad encoding is an error, as in the reference"""
    raw = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 96)
    out = np.zeros((raw.shape[0], 36), dtype=np.uint64)
    for i in range(raw.shape[0]):
        row = np.ascontiguousarray(raw[i])
        _raise(lib().b200_g2_from_compressed(_p(out[i]), _p(row)), errors=(BAD_INPUT,), what="FromCompressedG2")
    return out


def generate_testing_setup_g2(secret: int, n: int) -> np.ndarray:
    """setup.go:9-26 GenerateTestingSetup (G2 half): [secret^i * GenG2] for i < n (host: one scalar multiplication each)."""
    out = np.zeros((n, 36), dtype=np.uint64)
    s = fr_from_ints([secret % R_MOD])
    _raise(lib().b200_generate_testing_setup_g2(_p(s), n, _p(out)), what="GenerateTestingSetup (G2)")
    return out


def pairings_verify(a1, a2, b1, b2) -> bool:
    """bls/bls_kilic.go:152-158 PairingsVerify: e(a1, a2) == e(b1, b2)"""
    a1, b1 = _g1(a1)[0].copy(), _g1(b1)[0].copy()
    a2, b2 = _g2(a2)[0].copy(), _g2(b2)[0].copy()
    ok = C.c_int(0)
    _raise(lib().b200_pairings_verify(_p(a1), _p(a2), _p(b1), _p(b2), C.byref(ok)), what="PairingsVerify")
    return bool(ok.value)


def pairing(p, q) -> list:
    """e(p, q) as 12 integers a_0, b_0, .., a_5, b_5: sum (a_i + b_i u) w^i in Fp2[w] / (w^6 - (1 + u))"""
    p, q = _g1(p)[0].copy(), _g2(q)[0].copy()
    out = np.zeros((12, 6), dtype=np.uint64)
    _raise(lib().b200_pairing(_p(p), _p(q), _p(out)), what="pairing")
    return [sum(int(v) << (64 * j) for j, v in enumerate(row)) for row in out]


def lincomb_g1(points, scalars) -> np.ndarray:
    """bls/bls_kilic.go:132-150 LinCombG1 (device MSM); panics on a length mismatch."""
    pts, sc = _g1(points), _fr(scalars)
    if pts.shape[0] != sc.shape[0]:
        raise KZGPanic(LEN_MISMATCH, "got %d numbers and %d points" % (sc.shape[0], pts.shape[0]))
    out = np.zeros(18, dtype=np.uint64)
    _raise(lib().b200_g1_lincomb(_p(pts), _p(sc), pts.shape[0], _p(out)), what="LinCombG1")
    return out


def g1_mul_many(points, scalars) -> np.ndarray:
    pts, sc = _g1(points), _fr(scalars)
    assert pts.shape[0] == sc.shape[0]
    out = np.zeros_like(pts)
    _raise(lib().b200_g1_mul_many(_p(pts), _p(sc), pts.shape[0], _p(out)), what="MulG1 batch")
    return out


def generate_testing_setup_g1(secret: int, n: int) -> np.ndarray:
    """setup.go:9-26 GenerateTestingSetup (G1 half): [secret^i * G] for i < n, computed on the device."""
    out = np.zeros((n, 18), dtype=np.uint64)
    s = fr_from_ints([secret % R_MOD])
    _raise(lib().b200_generate_testing_setup_g1(_p(s), n, _p(out)), what="GenerateTestingSetup")
    return out


# ------------------------------------------------------------------------------- settings
class FFTSettings:
    """fft.go:34-61.  Holds the device-side domain tables."""

    def __init__(self, max_scale: int):
        h = C.c_void_p()
        _raise(lib().b200_fft_settings_new(max_scale, C.byref(h)), what="NewFFTSettings")
        self.h = h
        self.max_scale = max_scale
        self.max_width = 1 << max_scale

    def close(self):
        if getattr(self, "h", None):
            lib().b200_fft_settings_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def expanded_roots_of_unity(self, reverse: bool = False) -> np.ndarray:
        out = np.zeros((self.max_width + 1, 4), dtype=np.uint64)
        _raise(lib().b200_fs_roots(self.h, int(reverse), _p(out)))
        return out

    def fft(self, vals, inv: bool = False) -> np.ndarray:
        """fft_fr.go:55-74 FFT (zero-pads to the next power of two); error if too large."""
        v = _fr(vals)
        n = v.shape[0]
        npow = 1 if n == 0 else 1 << (n - 1).bit_length()
        out = np.zeros((npow, 4), dtype=np.uint64)
        _raise(lib().b200_fft_fr(self.h, _p(v), n, int(inv), _p(out)), errors=(TOO_LARGE, NOT_POW2), what="FFT")
        return out

    def inplace_fft(self, vals, inv: bool = False) -> np.ndarray:
        """fft_fr.go:76-105 InplaceFFT: no padding; error when the length is not a power of two (:81-83) or too
        large (:78-80); a zero length makes the reference divide by zero (panic)."""
        v = _fr(vals)
        out = np.zeros_like(v)
        _raise(lib().b200_inplace_fft_fr(self.h, _p(v), v.shape[0], int(inv), _p(out)), errors=(TOO_LARGE, NOT_POW2), what="InplaceFFT")
        return out

    def fft_batch(self, vals, inv: bool = False) -> np.ndarray:
        v = np.ascontiguousarray(vals, dtype=np.uint64)
        batch, n = v.shape[0], v.shape[1]
        npow = 1 if n == 0 else 1 << (n - 1).bit_length()
        out = np.zeros((batch, npow, 4), dtype=np.uint64)
        _raise(lib().b200_fft_fr_batch(self.h, _p(v), n, batch, int(inv), _p(out)), errors=(TOO_LARGE, NOT_POW2), what="FFT")
        return out

    def fft_g1(self, vals, inv: bool = False) -> np.ndarray:
        """fft_g1.go:58-94 FFTG1; error if too large or not a power of two."""
        v = _g1(vals)
        out = np.zeros_like(v)
        _raise(lib().b200_fft_g1(self.h, _p(v), v.shape[0], int(inv), _p(out)), errors=(TOO_LARGE, NOT_POW2), what="FFTG1")
        return out

    def toeplitz_part2(self, toeplitz_coeffs, x_ext_fft) -> np.ndarray:
        """fk20_single.go:59-77 (method of KZGSettings in the reference; it only needs the FFT settings)."""
        c, x = _fr(toeplitz_coeffs), _g1(x_ext_fft)
        out = np.zeros_like(x)
        _raise(lib().b200_toeplitz_part2(self.h, _p(c), c.shape[0], _p(x), x.shape[0], _p(out)), what="ToeplitzPart2")
        return out

    def toeplitz_part3(self, h_ext_fft) -> np.ndarray:
        """fk20_single.go:80-87: first half of the inverse FFTG1"""
        h = _g1(h_ext_fft)
        out = np.zeros((h.shape[0] // 2, 18), dtype=np.uint64)
        _raise(lib().b200_toeplitz_part3(self.h, _p(h), h.shape[0], _p(out)), what="ToeplitzPart3")
        return out

    def evaluate_poly_in_evaluation_form(self, polys, xs, reverse_bit_order: bool = False) -> np.ndarray:
        """bls/globals.go:106-153 for a batch: polys (batch, n, 4) evaluations on this settings' order-n domain
        (natural order, or reverse bit order like eth.DomainFr), xs (batch, 4) -> y (batch, 4)."""
        p = np.ascontiguousarray(polys, dtype=np.uint64)
        if p.ndim == 2:
            p = p[None]
        x = _fr(xs)
        assert x.shape[0] == p.shape[0]
        y = np.zeros((p.shape[0], 4), dtype=np.uint64)
        _raise(lib().b200_evaluate_poly_in_evaluation_form_batch(self.h, _p(p), _p(x), p.shape[1], p.shape[0], int(reverse_bit_order), _p(y)),
               what="EvaluatePolyInEvaluationForm")
        return y

    def fft_g1_batch(self, vals, inv: bool = False) -> np.ndarray:
        v = np.ascontiguousarray(vals, dtype=np.uint64)
        batch, n = v.shape[0], v.shape[1]
        out = np.zeros_like(v)
        _raise(lib().b200_fft_g1_batch(self.h, _p(v), n, batch, int(inv), _p(out)), errors=(TOO_LARGE, NOT_POW2), what="FFTG1")
        return out

    def das_fft_extension(self, vals) -> np.ndarray:
        """das_extension.go:71-84 DASFFTExtension (the reference works in place; a copy is returned)."""
        v = _fr(vals).copy()
        _raise(lib().b200_das_fft_extension(self.h, _p(v), v.shape[0]), what="DASFFTExtension")
        return v

    def das_fft_extension_g1(self, vals) -> np.ndarray:
        """das_extension.go:71-84 over G1 points (fk20_multi.go:96 TODO): odd-index evaluations from the even ones."""
        v = _g1(vals).copy()
        _raise(lib().b200_das_fft_extension_g1(self.h, _p(v), v.shape[0]), what="DASFFTExtension (G1)")
        return v

    def das_fft_extension_batch(self, vals) -> np.ndarray:
        v = np.ascontiguousarray(vals, dtype=np.uint64).copy()
        _raise(lib().b200_das_fft_extension_batch(self.h, _p(v), v.shape[1], v.shape[0]), what="DASFFTExtension")
        return v

    def zero_poly_via_multiplication(self, missing_indices, length: int):
        """zero_poly.go:116-217 -> (zeroEval, zeroPoly)"""
        m = np.ascontiguousarray(np.asarray(missing_indices, dtype=np.uint64))
        ze = np.zeros((length, 4), dtype=np.uint64)
        zp = np.zeros((length, 4), dtype=np.uint64)
        _raise(lib().b200_zero_poly_via_multiplication(self.h, _p(m), m.shape[0], length, _p(ze), _p(zp)),
               what="ZeroPolyViaMultiplication")
        return ze, zp

    def recover_poly_from_samples(self, samples, present) -> np.ndarray:
        """recover_from_samples.go:42-109; samples[i] is ignored where present[i] == 0 (nil)."""
        s = _fr(samples)
        pr = np.ascontiguousarray(present, dtype=np.uint8)
        out = np.zeros_like(s)
        _raise(lib().b200_recover_poly_from_samples(self.h, _p(s), _p(pr), s.shape[0], _p(out)),
               errors=(RECOVERY,), what="RecoverPolyFromSamples")
        return out

    def recover_poly_from_samples_batch(self, samples, present) -> np.ndarray:
        s = np.ascontiguousarray(samples, dtype=np.uint64)
        pr = np.ascontiguousarray(present, dtype=np.uint8)
        out = np.zeros_like(s)
        _raise(lib().b200_recover_poly_from_samples_batch(self.h, _p(s), _p(pr), s.shape[1], s.shape[0], _p(out)),
               errors=(RECOVERY,), what="RecoverPolyFromSamples")
        return out


class KZGSettings:
    """kzg.go:11-36.  secret_g2 (optional, (m, 36) uint64): the points CheckProofSingle / CheckProofMulti read; it may
    hold fewer entries than SecretG1 (the checks need SecretG2[1] and SecretG2[len(ys)]); secret_g2_len is the length
    NewKZGSettings compares with len(SecretG1) (kzg.go:22-24) and defaults to it."""

    def __init__(self, fs: FFTSettings, secret_g1, secret_g2_len=None, secret_g2=None):
        g = _g1(secret_g1)
        h = C.c_void_p()
        n2 = g.shape[0] if secret_g2_len is None else secret_g2_len
        _raise(lib().b200_kzg_settings_new(fs.h, _p(g), g.shape[0], n2, C.byref(h)), what="NewKZGSettings")
        self.h, self.fs = h, fs
        if secret_g2 is not None:
            q = _g2(secret_g2)
            _raise(lib().b200_kzg_settings_set_secret_g2(h, _p(q), q.shape[0]), what="NewKZGSettings (SecretG2)")
        self._g1_bytes = g if g.shape[0] <= (1 << 16) else None   # for setup_digest(); large setups are hashed at construction
        self._digest = None if self._g1_bytes is not None else _digest_points(g)

    def setup_digest(self) -> str:
        """SHA-256 of SecretG1 as handed in (ABI limbs) -- the key of the on-disk settings cache."""
        if self._digest is None:
            self._digest = _digest_points(self._g1_bytes)
        return self._digest

    def close(self):
        if getattr(self, "h", None):
            lib().b200_kzg_settings_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def commit_to_poly(self, coeffs) -> np.ndarray:
        """kzg_single_proofs.go:17-19"""
        c = _fr(coeffs)
        out = np.zeros(18, dtype=np.uint64)
        _raise(lib().b200_commit_to_poly(self.h, _p(c), c.shape[0], _p(out)), what="CommitToPoly")
        return out

    def blob_to_kzg_commitment_batch(self, blobs):
        """eth.BlobToKZGCommitment (eth/helpers.go:264-273 + :98-103) for blobs of n x 32 little-endian bytes;
        these settings must hold the bit-reversed Lagrange setup (eth/globals.go:48).
        Returns (commitments (batch, 48) uint8, ok (batch,) bool); ok is False where an element is >= r."""
        b = np.ascontiguousarray(blobs, dtype=np.uint8)
        batch, n = b.shape[0], b.shape[1]
        out = np.zeros((batch, 48), dtype=np.uint8)
        ok = np.zeros(batch, dtype=np.uint8)
        _raise(lib().b200_blob_to_kzg_commitment_batch(self.h, _p(b), n, batch, _p(out), _p(ok)), what="BlobToKZGCommitment")
        return out, ok.astype(bool)

    def compute_kzg_proof_batch(self, polys, zs):
        """eth.ComputeKZGProof (eth/helpers.go:179-203) per polynomial (evaluations in blob order) and challenge.
        Returns (proofs (batch, 48) uint8, y (batch, 4) uint64 canonical, ok (batch,) bool)."""
        p = np.ascontiguousarray(polys, dtype=np.uint64)
        z = np.ascontiguousarray(zs, dtype=np.uint64).reshape(-1, 4)
        batch, n = p.shape[0], p.shape[1]
        assert z.shape[0] == batch
        proofs = np.zeros((batch, 48), dtype=np.uint8)
        y = np.zeros((batch, 4), dtype=np.uint64)
        ok = np.zeros(batch, dtype=np.uint8)
        _raise(lib().b200_compute_kzg_proof_batch(self.h, _p(p), _p(z), n, batch, _p(proofs), _p(y), _p(ok)),
               errors=(TOO_LARGE, NOT_POW2, LEN_MISMATCH), what="ComputeKZGProof")
        return proofs, y, ok.astype(bool)

    def check_proof_multi_g1(self, commitments, xs, ys):
        """kzg_multi_proofs.go:47-88 CheckProofMulti without the G2 arithmetic and the pairing, for a batch of samples:
        ys (batch, n, 4).  Returns (commitment - [interpolation_polynomial(s)]_1 (batch, 18), x^n (batch, 4)); the
        caller checks e(first, [1]_2) == e(proof, SecretG2[n] - [x^n]_2)."""
        c, x = _g1(commitments), _fr(xs)
        y = np.ascontiguousarray(ys, dtype=np.uint64)
        if y.ndim == 2:
            y = y[None]
        batch, n = y.shape[0], y.shape[1]
        assert c.shape[0] == batch and x.shape[0] == batch
        out = np.zeros((batch, 18), dtype=np.uint64)
        xn = np.zeros((batch, 4), dtype=np.uint64)
        _raise(lib().b200_check_proof_multi_g1_batch(self.h, _p(c), _p(x), _p(y), n, batch, _p(out), _p(xn)), what="CheckProofMulti (G1 side)")
        return out, xn

    def check_proof_single(self, commitment, proof, x, y) -> bool:
        """kzg_single_proofs.go:57-75 CheckProofSingle (x, y: (4,) uint64 canonical)"""
        c, pr, xx, yy = _g1(commitment)[0].copy(), _g1(proof)[0].copy(), _fr(x)[0].copy(), _fr(y)[0].copy()
        ok = C.c_int(0)
        _raise(lib().b200_check_proof_single(self.h, _p(c), _p(pr), _p(xx), _p(yy), C.byref(ok)), what="CheckProofSingle")
        return bool(ok.value)

    def check_proof_single_batch(self, commitments, proofs, xs, ys) -> np.ndarray:
        """CheckProofSingle for a batch: the G1 sides in one device call, the pairing checks spread over the host cores."""
        c, pr, xx, yy = _g1(commitments), _g1(proofs), _fr(xs), _fr(ys)
        batch = c.shape[0]
        assert pr.shape[0] == batch and xx.shape[0] == batch and yy.shape[0] == batch
        ok = np.zeros(batch, dtype=np.uint8)
        _raise(lib().b200_check_proof_single_batch(self.h, _p(c), _p(pr), _p(xx), _p(yy), batch, _p(ok)), what="CheckProofSingle")
        return ok.astype(bool)

    def check_proof_multi(self, commitment, proof, x, ys) -> bool:
        """kzg_multi_proofs.go:47-88 CheckProofMulti (ys: (n, 4) uint64, n a power of two)"""
        c, pr, xx, yy = _g1(commitment)[0].copy(), _g1(proof)[0].copy(), _fr(x)[0].copy(), _fr(ys)
        ok = C.c_int(0)
        _raise(lib().b200_check_proof_multi(self.h, _p(c), _p(pr), _p(xx), _p(yy), yy.shape[0], C.byref(ok)), what="CheckProofMulti")
        return bool(ok.value)

    def check_proof_multi_batch(self, commitments, proofs, xs, ys) -> np.ndarray:
        """CheckProofMulti for a batch of samples: ys (batch, n, 4)"""
        c, pr, xx = _g1(commitments), _g1(proofs), _fr(xs)
        y = np.ascontiguousarray(ys, dtype=np.uint64)
        batch, n = y.shape[0], y.shape[1]
        assert c.shape[0] == batch and pr.shape[0] == batch and xx.shape[0] == batch
        ok = np.zeros(batch, dtype=np.uint8)
        _raise(lib().b200_check_proof_multi_batch(self.h, _p(c), _p(pr), _p(xx), _p(y), n, batch, _p(ok)), what="CheckProofMulti")
        return ok.astype(bool)

    @staticmethod
    def _random_scalars(batch: int) -> np.ndarray:
        """`batch` non-zero scalars below 2^248 (< r) from the OS CSPRNG, as (batch, 4) uint64 limbs"""
        import secrets
        raw = np.frombuffer(secrets.token_bytes(batch * 32), dtype=np.uint8).reshape(batch, 32).copy()
        raw[:, 31] = 0
        raw[:, 0] |= 1
        return raw.view(np.uint64).reshape(batch, 4)

    def check_proof_single_aggregate(self, commitments, proofs, xs, ys, rs=None) -> bool:
        """All proofs of the batch with one pairing (random linear combination; three device MSMs).  rs: non-zero
        unpredictable scalars, drawn from `secrets` when omitted."""
        c, pr, xx, yy = _g1(commitments), _g1(proofs), _fr(xs), _fr(ys)
        batch = c.shape[0]
        assert pr.shape[0] == batch and xx.shape[0] == batch and yy.shape[0] == batch
        r = self._random_scalars(batch) if rs is None else _fr(rs)
        ok = C.c_int(0)
        _raise(lib().b200_check_proof_single_aggregate(self.h, _p(c), _p(pr), _p(xx), _p(yy), _p(r), batch, C.byref(ok)),
               what="CheckProofSingle (aggregate)")
        return bool(ok.value)

    def check_proof_multi_aggregate(self, commitments, proofs, xs, ys, rs=None) -> bool:
        """CheckProofMulti for a batch of samples (ys (batch, n, 4)) with one pairing."""
        c, pr, xx = _g1(commitments), _g1(proofs), _fr(xs)
        y = np.ascontiguousarray(ys, dtype=np.uint64)
        batch, n = y.shape[0], y.shape[1]
        assert c.shape[0] == batch and pr.shape[0] == batch and xx.shape[0] == batch
        r = self._random_scalars(batch) if rs is None else _fr(rs)
        ok = C.c_int(0)
        _raise(lib().b200_check_proof_multi_aggregate(self.h, _p(c), _p(pr), _p(xx), _p(y), n, _p(r), batch, C.byref(ok)),
               what="CheckProofMulti (aggregate)")
        return bool(ok.value)

    def commit_to_poly_batch(self, coeffs) -> np.ndarray:
        c = np.ascontiguousarray(coeffs, dtype=np.uint64)
        out = np.zeros((c.shape[0], 18), dtype=np.uint64)
        _raise(lib().b200_commit_to_poly_batch(self.h, _p(c), c.shape[1], c.shape[0], _p(out)), what="CommitToPoly")
        return out


def _digest_points(g: np.ndarray) -> str:
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(g, dtype=np.uint64).tobytes()).hexdigest()


def _fk20_cached_handle(ks, n2: int, chunk_len: int, offsets, cache_dir: str, build):
    """On-disk cache of the xExtFFT files (kzg.go:57-62, 101-114), keyed by a digest of the setup, the FFT width, n2,
    the chunk length and the offsets.  build() -> handle computes them on the device; a cache hit adopts the stored
    files through b200_fk20_settings_new_from_x_ext_fft (the window tables are rebuilt either way)."""
    import hashlib
    lo, hi = (0, chunk_len) if offsets is None else (offsets.start, offsets.stop)
    key = hashlib.sha256(("%s|%d|%d|%d|%d|%d" % (ks.setup_digest(), ks.fs.max_scale, n2, chunk_len, lo, hi)).encode()).hexdigest()[:32]
    path = os.path.join(cache_dir, "fk20_xextfft_%s.npy" % key)
    k2 = n2 // chunk_len
    if os.path.exists(path):
        files = np.load(path)
        if files.shape == (hi - lo, k2, 18) and files.dtype == np.uint64:
            h = C.c_void_p()
            _raise(lib().b200_fk20_settings_new_from_x_ext_fft(ks.h, n2, chunk_len, lo, hi, _p(np.ascontiguousarray(files)), C.byref(h)),
                   what="FK20 settings from cache")
            return h, True
    h = build()
    os.makedirs(cache_dir, exist_ok=True)
    files = np.zeros((hi - lo, k2, 18), dtype=np.uint64)
    for f in range(lo, hi):
        _raise(lib().b200_fk20_x_ext_fft(h, f, _p(files[f - lo])))
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as fh:
        np.save(fh, files)
    os.replace(tmp, path)
    return h, False


class _FK20Base:
    def close(self):
        if getattr(self, "h", None):
            lib().b200_fk20_settings_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def x_ext_fft(self, file: int = 0) -> np.ndarray:
        out = np.zeros((self.n2 // self.chunk_len, 18), dtype=np.uint64)
        _raise(lib().b200_fk20_x_ext_fft(self.h, file, _p(out)))
        return out

    def last_launch_count(self) -> int:
        return int(lib().b200_fk20_last_launch_count(self.h))


class FK20SingleSettings(_FK20Base):
    """kzg.go:38-64"""

    def __init__(self, ks: KZGSettings, n2: int, cache_dir: str = None):
        """cache_dir: keep / reuse xExtFFT on disk (keyed by a digest of the setup); default: recompute, as kzg.go:43-64"""
        def build():
            h = C.c_void_p()
            _raise(lib().b200_fk20_single_settings_new(ks.h, n2, C.byref(h)), what="NewFK20SingleSettings")
            return h
        self.from_cache = False
        if cache_dir is None:
            h = build()
        else:
            h, self.from_cache = _fk20_cached_handle(ks, n2, 1, None, cache_dir, build)
        self.h, self.ks, self.n2, self.chunk_len = h, ks, n2, 1

    def fk20_single(self, poly) -> np.ndarray:
        """fk20_single.go:122-134: n proofs, natural order"""
        p = _fr(poly)
        out = np.zeros((p.shape[0], 18), dtype=np.uint64)
        _raise(lib().b200_fk20_single(self.h, _p(p), p.shape[0], _p(out)), what="FK20Single")
        return out

    def fk20_single_da_optimized(self, poly) -> np.ndarray:
        """fk20_single.go:139-172"""
        p = _fr(poly)
        out = np.zeros((p.shape[0], 18), dtype=np.uint64)
        _raise(lib().b200_fk20_single_da_optimized(self.h, _p(p), p.shape[0], _p(out)), what="FK20SingleDAOptimized")
        return out

    def da_using_fk20(self, poly) -> np.ndarray:
        """fk20_single.go:176-196: 2n proofs in reverse bit order"""
        p = _fr(poly)
        out = np.zeros((2 * p.shape[0], 18), dtype=np.uint64)
        _raise(lib().b200_da_using_fk20(self.h, _p(p), p.shape[0], _p(out)), what="DAUsingFK20")
        return out

    def da_using_fk20_batch(self, polys) -> np.ndarray:
        """DAUsingFK20 (fk20_single.go:176-196) for every polynomial of the batch: (batch, 2n, 18)"""
        p = np.ascontiguousarray(polys, dtype=np.uint64)
        batch, n = p.shape[0], p.shape[1]
        out = np.zeros((batch, 2 * n, 18), dtype=np.uint64)
        _raise(lib().b200_da_using_fk20_batch(self.h, _p(p), n, batch, _p(out)), what="DAUsingFK20 (batch)")
        return out

    def commit_fk20_batch(self, polys):
        """Headline unit: (CommitToPoly(p), FK20Single(p)) for every polynomial of the batch."""
        p = np.ascontiguousarray(polys, dtype=np.uint64)
        batch, n = p.shape[0], p.shape[1]
        commits = np.zeros((batch, 18), dtype=np.uint64)
        proofs = np.zeros((batch, n, 18), dtype=np.uint64)
        _raise(lib().b200_commit_fk20_batch(self.h, _p(p), n, batch, _p(commits), _p(proofs)), what="commit+FK20Single")
        return commits, proofs


    def commit_fk20_batch_compressed(self, polys):
        """commit_fk20_batch with 48-byte compressed outputs produced on the device."""
        p = np.ascontiguousarray(polys, dtype=np.uint64)
        batch, n = p.shape[0], p.shape[1]
        commits = np.zeros((batch, 48), dtype=np.uint8)
        proofs = np.zeros((batch, n, 48), dtype=np.uint8)
        _raise(lib().b200_commit_fk20_batch_compressed(self.h, _p(p), n, batch, _p(commits), _p(proofs)), what="commit+FK20Single (compressed)")
        return commits, proofs


class FK20MultiSettings(_FK20Base):
    """kzg.go:66-116"""

    def __init__(self, ks: KZGSettings, n2: int, chunk_len: int, offsets: range = None, cache_dir: str = None):
        """offsets: build and keep only the xExtFFT files of these chunk offsets (one rank of the offset-sharded
        multi-GPU form, go_kzg_b200/multi_gpu.py); default: all of them, as kzg.go:73-116 does.
        cache_dir: keep / reuse the files on disk (keyed by a digest of the setup)."""
        def build():
            h = C.c_void_p()
            if offsets is None:
                _raise(lib().b200_fk20_multi_settings_new(ks.h, n2, chunk_len, C.byref(h)), what="NewFK20MultiSettings")
            else:
                _raise(lib().b200_fk20_multi_settings_new_sharded(ks.h, n2, chunk_len, offsets.start, offsets.stop, C.byref(h)),
                       what="NewFK20MultiSettings (sharded)")
            return h
        self.from_cache = False
        if cache_dir is None or chunk_len < 1 or n2 < 2 or (n2 & (n2 - 1)):
            h = build()
        else:
            h, self.from_cache = _fk20_cached_handle(ks, n2, chunk_len, offsets, cache_dir, build)
        self.h, self.ks, self.n2, self.chunk_len, self.offsets = h, ks, n2, chunk_len, offsets

    def fk20_multi_da_optimized(self, poly) -> np.ndarray:
        """fk20_multi.go:58-109: n2 coefficients (upper half zero) -> 2k proofs"""
        p = _fr(poly)
        out = np.zeros((p.shape[0] // self.chunk_len, 18), dtype=np.uint64)
        _raise(lib().b200_fk20_multi_da_optimized(self.h, _p(p), p.shape[0], _p(out)), what="FK20MultiDAOptimized")
        return out

    def da_using_fk20_multi(self, poly) -> np.ndarray:
        """fk20_multi.go:113-133: 2k proofs in reverse bit order"""
        p = _fr(poly)
        out = np.zeros((2 * p.shape[0] // self.chunk_len, 18), dtype=np.uint64)
        _raise(lib().b200_da_using_fk20_multi(self.h, _p(p), p.shape[0], _p(out)), what="DAUsingFK20Multi")
        return out
