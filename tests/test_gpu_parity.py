"""Parity of the CUDA path against the oracle, through the C ABI (python mirror of the Go API).

Everything here needs a B200 (`-m gpu`).  G1 results are compared on their 48-byte compressed
encodings (Jacobian limbs are representation dependent; the reference itself only compares
projectively, bls/bls_kilic.go:106), Fr results on canonical limbs: bit-exact, no tolerance."""
import random

import numpy as np
import pytest

import go_kzg_b200 as kzg
from kzg_test_util import blob_polys, random_fr_ints
from oracle import cref, pyref

pytestmark = pytest.mark.gpu
R = pyref.R_MOD


def cmp_g1(got, want):
    assert np.array_equal(kzg.g1_to_compressed(got), cref.g1_compress(want))


# ------------------------------------------------------------------------------ primitives
def test_device_field_and_group_selftest():
    import ctypes as C
    bad = (C.c_uint64 * 17)()
    assert kzg.lib().b200_selftest_field(4096, 7, bad) == 0
    assert sum(bad) == 0, "per-check mismatch counts: " + " ".join(str(int(v)) for v in bad)


# ------------------------------------------------------------------------------ Fr FFT
def test_fft_fr_golden(goldens):
    """fft_fr_test.go:32-71 TestInvFFT"""
    g = goldens["inv_fft_scale4"]
    fs = kzg.FFTSettings(4)
    assert kzg.fr_to_ints(fs.fft(kzg.fr_from_ints(g["input"]), True)) == [int(v) for v in g["expected"]]


def test_fs_roots_match_oracle():
    fs, fo = kzg.FFTSettings(6), pyref.FFTSettings(6)
    assert kzg.fr_to_ints(fs.expanded_roots_of_unity()) == fo.expanded
    assert kzg.fr_to_ints(fs.expanded_roots_of_unity(True)) == fo.expanded[::-1]


@pytest.mark.parametrize("scale", [0, 1, 2, 3, 5, 8, 10, 12, 13, 14, 15])
def test_fft_fr_vs_oracle(scale):
    n = 1 << scale
    fs, fo = kzg.FFTSettings(max(scale, 4)), cref.FFTSettings(max(scale, 4))
    v = kzg.fr_from_ints(random_fr_ints(n, scale))
    for inv in (False, True):
        assert np.array_equal(fs.fft(v, inv), fo.fft(v, inv))
    back = fs.fft(fs.fft(v), True)                       # fft_fr_test.go:9-30 round trip
    assert np.array_equal(back, v)


def test_fft_fr_subsize_and_padding():
    """smaller transforms reuse the big domain with stride MaxWidth / n (fft_fr.go:89,100);
    non power-of-two inputs are zero padded (fft_fr.go:60-68)"""
    fs, fo = kzg.FFTSettings(14), cref.FFTSettings(14)
    for n in (1, 3, 16, 1000, 4097):
        v = kzg.fr_from_ints(random_fr_ints(n, n))
        for inv in (False, True):
            assert np.array_equal(fs.fft(v, inv), fo.fft(v, inv))
    with pytest.raises(kzg.KZGError):                    # fft_fr.go:57-59
        kzg.FFTSettings(3).fft(kzg.fr_from_ints(list(range(9))))


def test_fft_fr_batch():
    fs, fo = kzg.FFTSettings(13), cref.FFTSettings(13)
    v = np.stack([kzg.fr_from_ints(random_fr_ints(8192, 100 + b)) for b in range(3)])
    out = fs.fft_batch(v)
    for b in range(3):
        assert np.array_equal(out[b], fo.fft(v[b]))


# ------------------------------------------------------------------------------ DAS extension
def test_das_ext_golden(goldens):
    """das_extension_test.go:11-40"""
    g = goldens["das_ext_scale4"]
    fs = kzg.FFTSettings(4)
    assert kzg.fr_to_ints(fs.das_fft_extension(kzg.fr_from_ints(g["input"]))) == [int(v) for v in g["expected"]]


@pytest.mark.parametrize("scale", [2, 4, 7, 9, 13, 14, 15])
def test_das_ext_vs_oracle(scale):
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    even = kzg.fr_from_ints(random_fr_ints(1 << (scale - 1), scale))
    odd = fs.das_fft_extension(even)
    assert np.array_equal(odd, fo.das_fft_extension(even))
    # das_extension_test.go:42-86: the interleaved data has a zero upper half of coefficients
    inter = np.empty((1 << scale, 4), dtype=np.uint64)
    inter[0::2], inter[1::2] = even, odd
    coeffs = fs.fft(inter, True)
    assert not coeffs[(1 << scale) // 2:].any()


def test_das_ext_oversized_domain_and_panic():
    # MaxWidth > 2 n: the reference still indexes the full domain with stride 1 (das_extension.go:75)
    fs, fo = kzg.FFTSettings(8), cref.FFTSettings(8)
    v = kzg.fr_from_ints(random_fr_ints(32, 1))
    assert np.array_equal(fs.das_fft_extension(v), fo.das_fft_extension(v))
    with pytest.raises(kzg.KZGPanic):                    # das_extension.go:72-74
        kzg.FFTSettings(3).das_fft_extension(kzg.fr_from_ints(list(range(8))))


# ------------------------------------------------------------------------------ G1 FFT
@pytest.mark.parametrize("n", [1, 2, 4, 16, 64])
def test_fft_g1_vs_oracle(n):
    rng = random.Random(n)
    scale = max(n.bit_length() - 1, 1)
    fs, fo = kzg.FFTSettings(scale + 1), cref.FFTSettings(scale + 1)   # sub-size: stride 2
    ks = [rng.randrange(R) for _ in range(n)]
    if n >= 4:
        ks[1] = 0                   # infinity among the inputs
        ks[3] = ks[2]               # equal points -> doubling branch inside butterflies
    pts = cref.g1_mul_gen(ks)
    for inv in (False, True):
        cmp_g1(fs.fft_g1(pts, inv), fo.fft_g1(pts, inv))


def test_fft_g1_constant_and_zero_vectors():
    fs = kzg.FFTSettings(4)
    P = cref.g1_mul_gen([5])[0]
    same = np.stack([P] * 16)
    out = fs.fft_g1(same)           # DFT of a constant: (16 P, inf, inf, ...)
    want = np.zeros((16, 18), dtype=np.uint64)
    want[0] = cref.g1_mul_gen([80])[0]
    cmp_g1(out, want)
    cmp_g1(fs.fft_g1(np.zeros((16, 18), dtype=np.uint64)), np.zeros((16, 18), dtype=np.uint64))


def test_fft_g1_errors():
    """fft_g1.go:60-65"""
    fs = kzg.FFTSettings(3)
    with pytest.raises(kzg.KZGError):
        fs.fft_g1(np.zeros((16, 18), dtype=np.uint64))
    with pytest.raises(kzg.KZGError):
        fs.fft_g1(np.zeros((6, 18), dtype=np.uint64))


def test_fft_g1_lagrange_kat(trusted_setup_bytes):
    """eth/trusted_setup.json: FFTG1(setup_G1, inv) == setup_G1_lagrange, natural order, all
    4096 points -- the reference's shipped G1-IFFT known-answer vector."""
    s1, lag = trusted_setup_bytes
    fs = kzg.FFTSettings(12)
    out = fs.fft_g1(kzg.g1_from_compressed(s1), True)
    assert np.array_equal(kzg.g1_to_compressed(out), lag)


def test_fft_g1_batch_uses_shared_twiddle_programs():
    """batch of 32 transforms: lanes = blobs, width-5 NAF twiddle programs"""
    rng = random.Random(77)
    fs, fo = kzg.FFTSettings(4), pyref.FFTSettings(4)
    ks = [[rng.randrange(R) for _ in range(16)] for _ in range(32)]
    pts = np.stack([cref.g1_mul_gen(k) for k in ks])
    out = fs.fft_g1_batch(pts, False)
    for b in (0, 7, 31):
        cmp_g1(out[b], cref.g1_mul_gen(fo.fft(ks[b], False)))
    out = fs.fft_g1_batch(pts, True)
    for b in (0, 13, 31):
        cmp_g1(out[b], cref.g1_mul_gen(fo.fft(ks[b], True)))


@pytest.mark.parametrize("batch", [15, 16, 20, 33])
def test_fft_g1_batch_sizes_around_the_warp_padding(batch):
    """from 16 transforms on, the lanes of a butterfly are padded to whole warps and the sparse twiddle programs
    are used; below, per-lane fixed windows: both sides of the switch and ragged last warps"""
    rng = random.Random(batch)
    fs, fo = kzg.FFTSettings(3), pyref.FFTSettings(3)
    ks = [[rng.randrange(R) for _ in range(8)] for _ in range(batch)]
    pts = np.stack([cref.g1_mul_gen(k) for k in ks])
    for inv in (False, True):
        out = fs.fft_g1_batch(pts, inv)
        for b in (0, batch // 2, batch - 1):
            cmp_g1(out[b], cref.g1_mul_gen(fo.fft(ks[b], inv)))


# ------------------------------------------------------------------------------ LinCombG1 / commit
def test_empty_lincomb_is_infinity():
    """bls/bls_test.go:69-77 TestEmptyG1Lincomb"""
    out = kzg.lincomb_g1(np.zeros((0, 18), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64))
    assert not out.any()
    with pytest.raises(kzg.KZGPanic):                    # bls/bls_kilic.go:133-135
        kzg.lincomb_g1(np.zeros((2, 18), dtype=np.uint64), np.zeros((3, 4), dtype=np.uint64))


@pytest.mark.parametrize("n", [1, 5, 31, 32, 200])
def test_lincomb_vs_oracle(n):
    rng = random.Random(n)
    ks = [rng.randrange(R) for _ in range(n)]
    ss = [rng.randrange(R) for _ in range(n)]
    if n > 3:
        ss[1], ks[2], ss[3] = 0, 0, R - 1
    pts = cref.g1_mul_gen(ks)
    out = kzg.lincomb_g1(pts, kzg.fr_from_ints(ss))
    cmp_g1(out, cref.g1_mul_gen([sum(k * s for k, s in zip(ks, ss)) % R]))
    cmp_g1(out, cref.lincomb_g1(pts, cref.fr_to_limbs(ss)))


def test_mul_many_vs_oracle():
    rng = random.Random(3)
    ks = [rng.randrange(R) for _ in range(40)]
    ss = [0, 1, 2, R - 1, 1 << 128, (1 << 128) - 1] + [rng.randrange(R) for _ in range(34)]
    pts = cref.g1_mul_gen(ks)
    pts[5] = 0
    ks[5] = 0
    out = kzg.g1_mul_many(pts, kzg.fr_from_ints(ss))
    cmp_g1(out, cref.g1_mul_gen([k * s % R for k, s in zip(ks, ss)]))


def test_commit_to_poly_trusted_setup(trusted_setup_bytes):
    """config 2: CommitToPoly over eth/trusted_setup.json (secret 1337) == p(1337) G"""
    s1, _ = trusted_setup_bytes
    fs = kzg.FFTSettings(12)
    ks = kzg.KZGSettings(fs, kzg.g1_from_compressed(s1))
    coeffs = random_fr_ints(4096, 0xB2000000)
    com = ks.commit_to_poly(kzg.fr_from_ints(coeffs))
    cmp_g1(com, cref.g1_mul_gen([pyref.eval_poly(coeffs, 1337)]))
    short = ks.commit_to_poly(kzg.fr_from_ints(coeffs[:100]))      # SecretG1[:len(coeffs)]
    cmp_g1(short, cref.g1_mul_gen([pyref.eval_poly(coeffs[:100], 1337)]))


def test_kzg_settings_checks():
    """kzg.go:21-27 panics"""
    fs = kzg.FFTSettings(4)
    pts = cref.generate_setup_g1(5, 16)
    with pytest.raises(kzg.KZGPanic):
        kzg.KZGSettings(fs, pts, secret_g2_len=15)
    with pytest.raises(kzg.KZGPanic):
        kzg.KZGSettings(fs, pts[:8])


# ------------------------------------------------------------------------------ FK20
def test_fk20_single_reference_test_case(goldens):
    """fk20_single_test.go:11-44 (secret, polynomial and sizes of the reference's own test);
    every proof is checked, against the oracle and against the closed form."""
    t = goldens["fk20_single_test"]
    secret, poly, scale, n2 = int(t["secret"]), t["poly"], t["fft_scale"], t["n2"]
    n = len(poly)
    setup = cref.generate_setup_g1(secret, n2 + 1)
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, setup)
    fk = kzg.FK20SingleSettings(ks, n2)
    fo = cref.FK20(scale, setup, n2)
    cmp_g1(fk.x_ext_fft(), fo.x_ext_fft())
    p = kzg.fr_from_ints(poly)
    cmp_g1(ks.commit_to_poly(p), fo.commit(p))
    proofs = fk.fk20_single(p)
    cmp_g1(proofs, fo.fk20_single(p, da=False))
    cmp_g1(proofs, cref.g1_mul_gen(pyref.fk20_single_exponents(poly, secret)))
    da = fk.da_using_fk20(p)
    cmp_g1(da, fo.fk20_single(p, da=True))
    ext = np.concatenate([p, np.zeros_like(p)])
    dao = fk.fk20_single_da_optimized(ext)
    want = pyref.fk20_pipeline_exponents(pyref.FFTSettings(scale), poly, secret, True)
    pyref.reverse_bit_order(want)                        # DAUsingFK20 minus its final permutation
    cmp_g1(dao, cref.g1_mul_gen(want))
    ext[n + 1, 0] = 1
    with pytest.raises(kzg.KZGPanic):                    # fk20_single.go:150-154
        fk.fk20_single_da_optimized(ext)
    with pytest.raises(kzg.KZGPanic):                    # fk20_single.go:60-62
        fk.fk20_single(p[: n // 2])


def test_fk20_settings_panics():
    """kzg.go:44-52, 74-91"""
    fs = kzg.FFTSettings(4)
    ks = kzg.KZGSettings(fs, cref.generate_setup_g1(5, 16))
    for n2 in (32, 12, 1):
        with pytest.raises(kzg.KZGPanic):
            kzg.FK20SingleSettings(ks, n2)
    for n2, l in ((16, 16), (16, 3), (16, 0)):
        with pytest.raises(kzg.KZGPanic):
            kzg.FK20MultiSettings(ks, n2, l)


def test_fk20_multi_reference_test_case(goldens):
    """fk20_multi_test.go:11-91: chunk 16 x 32 chunks, the reference's secret; all 64 coset
    proofs against the oracle pipeline restated in the exponent."""
    t = goldens["fk20_multi_test"]
    secret, l, cc = int(t["secret"]), t["chunk_len"], t["chunk_count"]
    n = l * cc
    scale = (2 * n).bit_length() - 1
    poly = random_fr_ints(n, 11)
    setup = cref.generate_setup_g1(secret, 2 * n)
    fs = kzg.FFTSettings(scale)
    fk = kzg.FK20MultiSettings(kzg.KZGSettings(fs, setup), 2 * n, l)
    got = fk.da_using_fk20_multi(kzg.fr_from_ints(poly))
    want = pyref.fk20_multi_da_exponents(pyref.FFTSettings(scale), poly, secret, l)
    cmp_g1(got, cref.g1_mul_gen(want))
    fo = cref.FK20(scale, setup, 2 * n, l)
    cmp_g1(fk.x_ext_fft(3), fo.x_ext_fft(3))
    ext = kzg.fr_from_ints(poly + [0] * n)
    nat = fk.fk20_multi_da_optimized(ext)
    k2 = 2 * cc
    rev = [pyref.reverse_bits_limited(k2, i) for i in range(k2)]
    cmp_g1(nat[rev], cref.g1_mul_gen(want))


def _setup_8192(trusted_setup_bytes):
    s1, _ = trusted_setup_bytes
    first = kzg.g1_from_compressed(s1)
    rest = cref.g1_mul_gen([pow(1337, i, R) for i in range(4096, 8192)])
    return np.concatenate([first, rest])


def test_commit_fk20_batch_n4096(trusted_setup_bytes):
    """configs 2 + 3 at full size: 32 blobs of 4096 coefficients over the trusted setup
    (secret 1337, extended to 8192 points): commitments == p(s) G and all 4096 proofs of
    sampled blobs == (p(s) - p(w^i)) / (s - w^i) G."""
    fs = kzg.FFTSettings(13)
    fk = kzg.FK20SingleSettings(kzg.KZGSettings(fs, _setup_8192(trusted_setup_bytes)), 8192)
    batch = 32
    polys = blob_polys(batch, 4096)
    commits, proofs = fk.commit_fk20_batch(polys)
    assert fk.last_launch_count() > 0
    ints = [kzg.fr_to_ints(polys[b]) for b in range(batch)]
    cmp_g1(commits, cref.g1_mul_gen([pyref.eval_poly(p, 1337) for p in ints]))
    for b in (0, 17, 31):
        cmp_g1(proofs[b], cref.g1_mul_gen(pyref.fk20_single_exponents(ints[b], 1337)))
    # linearity (size independent): proofs(p0 + p1) == proofs(p0) + proofs(p1) at a few positions
    psum = kzg.fr_from_ints([(a + b) % R for a, b in zip(ints[0], ints[1])])
    single = fk.fk20_single(psum)           # batch of one: per-lane twiddle programs
    L = kzg.lib()
    for i in (0, 1, 2047, 4095):
        s = np.zeros(18, dtype=np.uint64)
        L.b200_g1_add(s.ctypes.data, proofs[0][i].ctypes.data, proofs[1][i].ctypes.data)
        assert L.b200_g1_equal(s.ctypes.data, single[i].ctypes.data) == 1


# ------------------------------------------------------------------------------ compressed outputs, eth blobs
def _brp(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2)


@pytest.mark.parametrize("n", [1, 15, 16, 17, 250])
def test_g1_compress_on_device(n, goldens):
    """ToCompressedG1 (bls/bls_kilic.go:114) normalised and compressed on the GPU == the oracle's and the
    host level-1 bytes: Jacobian inputs with Z != 1, infinities, group sizes around the 16-point
    inversion batches; the reference's golden vector (bls/bls_test.go:13,18)."""
    rnd = random.Random(n)
    gen = cref.g1_generator()
    pts = kzg.g1_mul_many(np.repeat(gen[None, :], n, axis=0), kzg.fr_from_ints([rnd.randrange(1, R) for _ in range(n)]))
    gold = goldens["point_compression"]                     # bls/bls_test.go:13,18 TestPointCompression
    if n > 2:
        pts[n // 2] = 0                                     # infinity (Z == 0)
        pts[n - 1] = gen                                    # Z == 1
        pts[0] = kzg.g1_mul_many(gen[None, :], kzg.fr_from_ints([int(gold["scalar"])]))[0]
    got = kzg.g1_to_compressed_device(pts)
    assert np.array_equal(got, cref.g1_compress(pts))
    assert np.array_equal(got, kzg.g1_to_compressed(pts))
    if n > 2:
        assert got[n // 2, 0] == 0xC0 and not got[n // 2, 1:].any()
        assert list(got[0]) == gold["expected"]


def test_commit_fk20_batch_compressed_matches_uncompressed(trusted_setup_bytes):
    """The compressed-output form of the headline unit returns exactly ToCompressedG1 of the Jacobian results."""
    fs = kzg.FFTSettings(13)
    fk = kzg.FK20SingleSettings(kzg.KZGSettings(fs, _setup_8192(trusted_setup_bytes)), 8192)
    polys = blob_polys(32, 4096, first_blob=500)
    commits, proofs = fk.commit_fk20_batch(polys)
    c48, p48 = fk.commit_fk20_batch_compressed(polys)
    assert np.array_equal(c48, cref.g1_compress(commits))
    for b in (0, 13, 31):
        assert np.array_equal(p48[b], cref.g1_compress(proofs[b]))
    assert np.array_equal(p48.reshape(-1, 48)[::97], kzg.g1_to_compressed(proofs.reshape(-1, 18)[::97]))


def test_blob_to_kzg_commitment(trusted_setup_bytes):
    """eth.BlobToKZGCommitment (eth/helpers.go:98-103, :264-273) over the bit-reversed Lagrange setup
    (eth/globals.go:48): for evaluations v of a polynomial c on the natural domain, the blob is brp(v) and the
    commitment is c(1337) G (known-secret closed form) == the oracle's LinCombG1 on the same inputs.
    A field element >= r makes BlobToPolynomial fail: ok = False."""
    s1, lag = trusted_setup_bytes
    n, bits = 4096, 12
    lagrange = kzg.g1_from_compressed(lag)
    perm = np.array([_brp(i, bits) for i in range(n)])
    fs = kzg.FFTSettings(bits)
    ks = kzg.KZGSettings(fs, lagrange[perm])
    batch = 5
    blobs = np.zeros((batch, n, 32), dtype=np.uint8)
    want = []
    ofs = cref.FFTSettings(bits)
    for b in range(batch):
        v = random_fr_ints(n, 0xE7000000 + b)
        coeffs = cref.limbs_to_fr(ofs.fft(cref.fr_to_limbs(v), True))
        want.append(pyref.eval_poly(coeffs, 1337))
        blob_vals = [v[perm[i]] for i in range(n)]
        blobs[b] = np.frombuffer(b"".join(x.to_bytes(32, "little") for x in blob_vals), dtype=np.uint8).reshape(n, 32)
    got, ok = ks.blob_to_kzg_commitment_batch(blobs)
    assert ok.all()
    assert np.array_equal(got, cref.g1_compress(cref.g1_mul_gen(want)))
    lin = cref.lincomb_g1(lagrange[perm], np.ascontiguousarray(blobs[0]).view(np.uint64).reshape(n, 4))
    assert np.array_equal(got[0], cref.g1_compress(lin[None, :])[0])
    bad = blobs.copy()
    bad[3, 77] = np.frombuffer(R.to_bytes(32, "little"), dtype=np.uint8)          # == r: not a field element
    got2, ok2 = ks.blob_to_kzg_commitment_batch(bad)
    assert list(ok2) == [True, True, True, False, True]
    assert np.array_equal(got2[[0, 1, 2, 4]], got[[0, 1, 2, 4]]) and not got2[3].any()


def test_compute_kzg_proof_evaluation_form(trusted_setup_bytes):
    """eth.ComputeKZGProof (eth/helpers.go:179-203) + EvaluatePolyInEvaluationForm (bls/globals.go:106-153):
    y == p(z) for the interpolating polynomial p (oracle inverse NTT + Horner, independent of the barycentric
    formula) and proof == ((p(s) - y) / (s - z)) G for the known secret s = 1337; z inside the domain is the
    reference's "invalid z challenge" error."""
    s1, lag = trusted_setup_bytes
    n, bits = 4096, 12
    perm = np.array([_brp(i, bits) for i in range(n)])
    fs = kzg.FFTSettings(bits)
    ks = kzg.KZGSettings(fs, kzg.g1_from_compressed(lag)[perm])
    ofs = cref.FFTSettings(bits)
    batch = 4
    polys = np.zeros((batch, n, 4), dtype=np.uint64)
    zs, want_y, want_q = [], [], []
    w = pow(7, (R - 1) // n, R)
    for b in range(batch):
        v = random_fr_ints(n, 0xE8000000 + b)
        coeffs = cref.limbs_to_fr(ofs.fft(cref.fr_to_limbs(v), True))
        z = random_fr_ints(1, 0xE9000000 + b)[0] if b != 2 else pow(w, int(perm[5]), R)      # blob 2: z == D[5]
        zs.append(z)
        y = pyref.eval_poly(coeffs, z)
        want_y.append(y)
        want_q.append((pyref.eval_poly(coeffs, 1337) - y) * pow((1337 - z) % R, -1, R) % R if b != 2 else 0)
        polys[b] = kzg.fr_from_ints([v[perm[i]] for i in range(n)])
    proofs, y, ok = ks.compute_kzg_proof_batch(polys, kzg.fr_from_ints(zs))
    assert list(ok) == [True, True, False, True]
    good = [0, 1, 3]
    assert [kzg.fr_to_ints(y[b:b + 1])[0] for b in good] == [want_y[b] for b in good]
    assert np.array_equal(proofs[good], cref.g1_compress(cref.g1_mul_gen([want_q[b] for b in good])))
    assert not proofs[2].any()
    with pytest.raises(kzg.KZGError):
        ks.compute_kzg_proof_batch(polys[:, :100], kzg.fr_from_ints(zs))               # not a power of two


# ------------------------------------------------------------------------------ zero poly / recovery
def test_zero_poly_golden(goldens):
    """zero_poly_test.go:133-198 TestFFTSettings_ZeroPolyViaMultiplication_Python"""
    g = goldens["zero_poly_scale4"]
    missing = [i for i, e in enumerate(g["exists"]) if not e]
    ze, zp = kzg.FFTSettings(4).zero_poly_via_multiplication(missing, 16)
    assert kzg.fr_to_ints(ze) == [int(v) for v in g["expected_eval"]]
    assert kzg.fr_to_ints(zp) == [int(v) for v in g["expected_poly"]]


@pytest.mark.parametrize("scale,nmiss", [(3, 1), (5, 7), (7, 63), (7, 64), (8, 100), (10, 512), (11, 700), (12, 2048), (14, 8192)])
def test_zero_poly_vs_oracle(scale, nmiss):
    """zero_poly_test.go:251-261 shape; single-leaf, multi-leaf and multi-round cases"""
    n = 1 << scale
    rng = random.Random(scale * 1000 + nmiss)
    missing = sorted(rng.sample(range(n), nmiss))
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    ze, zp = fs.zero_poly_via_multiplication(missing, n)
    ze_o, zp_o = fo.zero_poly(missing, n)
    assert np.array_equal(ze, ze_o) and np.array_equal(zp, zp_o)
    zi = kzg.fr_to_ints(ze)
    assert all((zi[i] == 0) == (i in set(missing)) for i in range(n))      # vanishes exactly on the missing set


def test_zero_poly_edge_cases():
    fs = kzg.FFTSettings(6)
    ze, zp = fs.zero_poly_via_multiplication([], 64)                        # zero_poly.go:117-119
    assert not ze.any() and not zp.any()
    with pytest.raises(kzg.KZGPanic):                                       # zero_poly.go:120-122
        fs.zero_poly_via_multiplication([1], 128)
    with pytest.raises(kzg.KZGPanic):                                       # zero_poly.go:123-125
        fs.zero_poly_via_multiplication([1], 48)
    with pytest.raises(kzg.KZGPanic):                                       # degree does not fit (zero_poly.go:133 / 207-209)
        fs.zero_poly_via_multiplication(list(range(64)), 64)
    # sub-domain with stride (MaxWidth / length = 4)
    fo = cref.FFTSettings(6)
    ze, zp = fs.zero_poly_via_multiplication([0, 3, 5], 16)
    ze_o, zp_o = fo.zero_poly([0, 3, 5], 16)
    assert np.array_equal(ze, ze_o) and np.array_equal(zp, zp_o)


def _recovery_case(scale, known_ratio, seed):
    """recover_from_samples_test.go:61-137: poly = [0 .. n/2-1, 0 ...], random erasures"""
    n = 1 << scale
    fo = pyref.FFTSettings(scale) if scale <= 10 else None
    poly = list(range(n // 2)) + [0] * (n // 2)
    data = (kzg.fr_to_ints(cref.FFTSettings(scale).fft(cref.fr_to_limbs(poly))) if fo is None else fo.fft(poly))
    rng = random.Random(seed)
    nmiss = n - int(n * known_ratio)
    miss = set(rng.sample(range(n), nmiss))
    present = np.array([0 if i in miss else 1 for i in range(n)], dtype=np.uint8)
    samples = kzg.fr_from_ints([0 if i in miss else d for i, d in enumerate(data)])
    return data, samples, present


def test_recover_simple():
    """recover_from_samples_test.go:10-59: n = 4, keep indices 0 and 3"""
    fs = kzg.FFTSettings(2)
    data = pyref.FFTSettings(2).fft([1, 2, 0, 0])
    rec = fs.recover_poly_from_samples(kzg.fr_from_ints([data[0], 0, 0, data[3]]), [1, 0, 0, 1])
    assert kzg.fr_to_ints(rec) == data


@pytest.mark.parametrize("scale,ratio,seed", [(4, 0.5, 0), (8, 0.7, 1), (10, 0.5, 2), (10, 0.95, 3), (14, 0.5, 14)])
def test_recover_vs_oracle(scale, ratio, seed):
    data, samples, present = _recovery_case(scale, ratio, seed)
    fs = kzg.FFTSettings(scale)
    rec = fs.recover_poly_from_samples(samples, present)
    assert kzg.fr_to_ints(rec) == data
    if scale <= 10:
        assert np.array_equal(rec, cref.FFTSettings(scale).recover(samples, present))


def test_recover_batch_and_errors():
    scale = 9
    fs = kzg.FFTSettings(scale)
    cases = [_recovery_case(scale, 0.5 + 0.1 * b, 50 + b) for b in range(4)]
    out = fs.recover_poly_from_samples_batch(np.stack([c[1] for c in cases]), np.stack([c[2] for c in cases]))
    for b in range(4):
        assert kzg.fr_to_ints(out[b]) == cases[b][0]
    data, samples, present = cases[2]
    # A corrupted known sample is NOT an error: any n - missing values are interpolated exactly by a
    # polynomial of degree < n - missing, so the division by the zero polynomial stays exact and the
    # final comparison (recover_from_samples.go:103-107) holds; device and oracle agree on the result.
    noisy = samples.copy()
    noisy[np.flatnonzero(present)[0], 0] ^= 1
    rec = fs.recover_poly_from_samples(noisy, present)
    assert np.array_equal(rec, cref.FFTSettings(scale).recover(noisy, present))
    assert np.array_equal(rec[present == 1], noisy[present == 1])
    with pytest.raises(kzg.KZGPanic):                                       # nothing missing: "bad zero eval" (:54-58)
        fs.recover_poly_from_samples(samples, np.ones(1 << scale, dtype=np.uint8))


@pytest.mark.parametrize("batch", [16, 18, 37])
def test_larger_batches_of_recovery_and_extension(batch):
    """Batches with a different missing ratio per polynomial: the index lists are compacted on the device from the presence
    masks (ragged counts inside one launch); every output against the data / its own single call."""
    scale = 8
    fs = kzg.FFTSettings(scale)
    cases = [_recovery_case(scale, 0.5 + 0.02 * (b % 20), 900 + b) for b in range(batch)]
    out = fs.recover_poly_from_samples_batch(np.stack([c[1] for c in cases]), np.stack([c[2] for c in cases]))
    for b in range(batch):
        assert kzg.fr_to_ints(out[b]) == cases[b][0]
    # the extension: batch == one call per polynomial
    evens = np.stack([kzg.fr_from_ints(random_fr_ints(1 << (scale - 1), 7000 + b)) for b in range(batch)])
    odds = fs.das_fft_extension_batch(evens)
    for b in (0, 1, batch // 2, batch - 1):
        assert np.array_equal(odds[b], fs.das_fft_extension(evens[b]))
    # a polynomial with nothing missing at the end of the batch: the whole call fails the way the single call does (:54-58)
    pres = np.stack([c[2] for c in cases])
    pres[batch - 1] = 1
    with pytest.raises(kzg.KZGPanic):
        fs.recover_poly_from_samples_batch(np.stack([c[1] for c in cases]), pres)


def test_das_extension_then_recovery_round_trip():
    """config 4 shape: extend 8192 -> 16384 samples, erase half, recover everything (size-independent property)"""
    scale = 14
    fs = kzg.FFTSettings(scale)
    even = kzg.fr_from_ints(random_fr_ints(1 << (scale - 1), 4242))
    odd = fs.das_fft_extension(even)
    full = np.empty((1 << scale, 4), dtype=np.uint64)
    full[0::2], full[1::2] = even, odd
    rng = random.Random(14)
    perm = list(range(1 << scale))
    rng.shuffle(perm)
    present = np.ones(1 << scale, dtype=np.uint8)
    present[perm[: 1 << (scale - 1)]] = 0
    samples = full.copy()
    samples[present == 0] = 0
    rec = fs.recover_poly_from_samples(samples, present)
    assert np.array_equal(rec, full)


# ------------------------------------------------------------------------------ multi-GPU building blocks
def test_generate_testing_setup_vs_oracle():
    """setup.go:9-26"""
    secret = 1927409816240961209460912649124
    cmp_g1(kzg.generate_testing_setup_g1(secret, 70), cref.generate_setup_g1(secret, 70))


def test_sharded_fk20_multi_blocks_on_one_gpu():
    """offset-sharded FK20 multi (config 5 layout) with the ranks played one after another on a
    single GPU: partials per offset range -> G1 sum -> finish == DAUsingFK20Multi"""
    import ctypes as C
    import torch
    from go_kzg_b200 import multi_gpu
    secret, l, cc = 1927409816240961209460912649124, 16, 32
    n = l * cc
    scale = (2 * n).bit_length() - 1
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, kzg.generate_testing_setup_g1(secret, 2 * n))
    fk = kzg.FK20MultiSettings(ks, 2 * n, l)
    poly = kzg.fr_from_ints(random_fr_ints(n, 21))
    want = fk.da_using_fk20_multi(poly)
    L, k2, world = kzg.lib(), 2 * cc, 3                          # 3 ranks: uneven offset ranges (5, 5, 6)
    d_poly = torch.from_numpy(poly.view(np.int64)).cuda()
    parts = torch.zeros((world, k2, 18), dtype=torch.int64, device="cuda")
    covered = []
    for r in range(world):
        rng_ = multi_gpu.offset_range(r, world, l)
        covered += list(rng_)
        assert L.b200_fk20_multi_partial_dev(fk.h, d_poly.data_ptr(), n, rng_.start, rng_.stop, parts[r].data_ptr(), None) == 0
    assert covered == list(range(l))
    d_sum = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    assert L.b200_g1_sum_dev(parts.data_ptr(), world, k2, d_sum.data_ptr(), None) == 0
    d_out = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    assert L.b200_fk20_multi_finish_dev(fk.h, d_sum.data_ptr(), 1, d_out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    cmp_g1(d_out.cpu().numpy().view(np.uint64), want)
    # block-sharded finish (the two G1 transforms spread over 2^s ranks), ranks played one after another
    for w2 in (1, 2, 4, 8):
        blocks = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
        for r in range(w2):
            blk = blocks[r * (k2 // w2):(r + 1) * (k2 // w2)]
            assert L.b200_fk20_multi_finish_local_dev(fk.h, d_sum.data_ptr(), r, w2, blk.data_ptr(), None) == 0
        d_out2 = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
        assert L.b200_fk20_multi_finish_merge_dev(fk.h, blocks.data_ptr(), w2, 1, d_out2.data_ptr(), None) == 0
        torch.cuda.synchronize()
        cmp_g1(d_out2.cpu().numpy().view(np.uint64), want)
        # the merge sharded as well: every rank finishes its positions of every block, parts gathered in rank order
        if multi_gpu.merge_sharded(w2, k2) or w2 == 1:
            parts2 = torch.zeros((w2, k2 // w2, 18), dtype=torch.int64, device="cuda")
            for r in range(w2):
                mine = blocks.clone()                                     # the rank's own copy of the all-gathered blocks
                assert L.b200_fk20_multi_finish_merge_part_dev(fk.h, mine.data_ptr(), r, w2, parts2[r].data_ptr(), None) == 0
            d_out3 = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
            assert L.b200_fk20_multi_finish_assemble_dev(fk.h, parts2.data_ptr(), w2, 1, d_out3.data_ptr(), None) == 0
            torch.cuda.synchronize()
            cmp_g1(d_out3.cpu().numpy().view(np.uint64), want)
    assert L.b200_fk20_multi_finish_local_dev(fk.h, d_sum.data_ptr(), 0, 3, blocks.data_ptr(), None) != 0      # not a power of two
    # the single-process entry of the sharded driver, and the point-range sharded commitment
    cmp_g1(multi_gpu.da_using_fk20_multi_sharded(fk, poly), want)
    c_parts = torch.zeros((world, 1, 18), dtype=torch.int64, device="cuda")
    for r in range(world):
        pr = multi_gpu.point_range(r, world, n)
        assert L.b200_commit_partial_dev(ks.h, d_poly.data_ptr(), pr.start, pr.stop, c_parts[r].data_ptr(), None) == 0
    c_sum = torch.zeros((1, 18), dtype=torch.int64, device="cuda")
    assert L.b200_g1_sum_dev(c_parts.data_ptr(), world, 1, c_sum.data_ptr(), None) == 0
    torch.cuda.synchronize()
    cmp_g1(c_sum.cpu().numpy().view(np.uint64), ks.commit_to_poly(poly).reshape(1, 18))
    del C


def test_fk20_multi_config5_full_size():
    """config 5 at full size on one GPU: n = 2^20 coefficients, chunk 16 -> 131072 coset proofs.
    Expected values come from the pipeline restated in the exponent (secret known) with the
    oracle's Fr FFT; a spread of positions is compared on compressed bytes, and one position
    against the closed form q(s) G with p = q (X^l - x^l) + r."""
    secret, l = 1927409816240961209460912649124, 16
    n = 1 << 20
    k = n // l
    k2, scale = 2 * k, 21
    fs = kzg.FFTSettings(scale)
    setup = kzg.generate_testing_setup_g1(secret, 1 << scale)
    ks = kzg.KZGSettings(fs, setup)
    fk = kzg.FK20MultiSettings(ks, 1 << scale, l)
    del setup
    from go_kzg_b200.synth import random_fr_limbs
    poly = random_fr_limbs(n, 5)
    got = fk.da_using_fk20_multi(poly)
    # --- exponent-domain restatement (pyref.fk20_multi_da_exponents) on the oracle's C Fr FFT
    fo = cref.FFTSettings(scale)                      # transforms of size 2k use stride 2^21 / 2k
    p_int = kzg.fr_to_ints(poly)
    spow = [1] * n
    for i in range(1, n):
        spow[i] = spow[i - 1] * secret % R
    h_ext = [0] * k2
    for off in range(l):
        start = n - l - 1 - off
        x = [spow[start - i * l] for i in range(k - 1)] + [0] * (k + 1)
        xf = kzg.fr_to_ints(fo.fft(cref.fr_to_limbs(x)))
        c = pyref.toeplitz_coeffs_step_strided(p_int, off, l)
        cf = kzg.fr_to_ints(fo.fft(cref.fr_to_limbs(c)))
        for j in range(k2):
            h_ext[j] = (h_ext[j] + cf[j] * xf[j]) % R
    h = kzg.fr_to_ints(fo.fft(cref.fr_to_limbs(h_ext), True))[:k] + [0] * k
    out = kzg.fr_to_ints(fo.fft(cref.fr_to_limbs(h)))
    pyref.reverse_bit_order(out)
    idx = list(range(0, k2, k2 // 61)) + [1, k2 - 1]
    cmp_g1(got[idx], cref.g1_mul_gen([out[i] for i in idx]))
    # closed form at one position
    pos = 12345
    w = pyref.scale2_root_of_unity(scale)
    xq = pow(w, pyref.reverse_bits_limited(k2, pos), R)          # x = w_2n^brp(pos): coset x <w_l> (fk20_multi_test.go:60-64)
    xl = pow(xq, l, R)
    rem = list(p_int)
    acc = 0
    for i in range(n - 1, l - 1, -1):                 # synthetic division by X^l - x^l, quotient evaluated at s on the fly
        acc = (acc + rem[i] * spow[i - l]) % R
        rem[i - l] = (rem[i - l] + rem[i] * xl) % R
    assert out[pos] == acc
