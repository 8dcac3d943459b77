"""Parity at the shapes bench.py times and BASELINE.json names (through the C ABI, `-m gpu`):
the headline call at batch 128, DAUsingFK20 / FK20SingleDAOptimized at n = 4096, generic LinCombG1 (Pippenger
bucket MSM) at n = 4096 and 65536 over arbitrary points, forward FFTG1 at 4096 and 8192, config 4 at n = 2^14
with batch 8.  Expected values: the exponent-domain oracle (secret known, every G1 output is k G with k from the
oracle's Fr arithmetic) and the closed form proof_i = (p(s) - p(x_i)) / (s - x_i) G of fk20_single.go:122-196."""
import random

import numpy as np
import pytest

import go_kzg_b200 as kzg
from kzg_test_util import blob_polys, random_fr_ints
from oracle import cref, pyref

pytestmark = pytest.mark.gpu
R = pyref.R_MOD
SECRET = 1337


def cmp_g1(got, want):
    assert np.array_equal(kzg.g1_to_compressed(got), cref.g1_compress(want))


def single_proof_exponents(poly, secret, out_scale):
    """Discrete logs of the single-point proofs at x_i = w^i, w of order 2^out_scale, natural order:
    (p(s) - p(x_i)) / (s - x_i); p(x_i) for all i from one oracle Fr FFT."""
    m = 1 << out_scale
    fo = cref.FFTSettings(out_scale)
    evals = cref.limbs_to_fr(fo.fft(cref.fr_to_limbs(list(poly) + [0] * (m - len(poly)))))
    ps = pyref.eval_poly(poly, secret)
    w = pyref.scale2_root_of_unity(out_scale)
    out, x = [], 1
    for i in range(m):
        out.append((ps - evals[i]) * pow((secret - x) % R, -1, R) % R)
        x = x * w % R
    return out


@pytest.fixture(scope="module")
def fk4096(trusted_setup_bytes):
    s1, _ = trusted_setup_bytes
    first = kzg.g1_from_compressed(s1)
    rest = cref.g1_mul_gen([pow(SECRET, i, R) for i in range(4096, 8192)])
    fs = kzg.FFTSettings(13)
    return kzg.FK20SingleSettings(kzg.KZGSettings(fs, np.concatenate([first, rest])), 8192)


def test_headline_call_at_the_benchmarked_batch(fk4096):
    """bench.py's step: b200_commit_fk20_batch with 128 blobs of 4096 coefficients (the whole-warp lane mapping
    with the batch a multiple of 32).  Every commitment, and all 4096 proofs of three blobs, against the closed
    form; one position of every blob as well."""
    batch = 128
    polys = blob_polys(batch, 4096)
    commits, proofs = fk4096.commit_fk20_batch(polys)
    ints = [kzg.fr_to_ints(polys[b]) for b in range(batch)]
    cmp_g1(commits, cref.g1_mul_gen([pyref.eval_poly(p, SECRET) for p in ints]))
    for b in (0, 77, 127):
        cmp_g1(proofs[b], cref.g1_mul_gen(single_proof_exponents(ints[b], SECRET, 12)))
    w = pyref.scale2_root_of_unity(12)
    pos = 2741
    x = pow(w, pos, R)
    want = [(pyref.eval_poly(p, SECRET) - pyref.eval_poly(p, x)) * pow((SECRET - x) % R, -1, R) % R for p in ints]
    cmp_g1(proofs[:, pos], cref.g1_mul_gen(want))


@pytest.mark.parametrize("batch", [1, 40])
def test_headline_call_ragged_batches(fk4096, batch):
    """batch 1 (per-lane twiddle programs) and 40 (whole-warp lanes with a ragged last warp) at n = 4096"""
    polys = blob_polys(batch, 4096, first_blob=9000)
    commits, proofs = fk4096.commit_fk20_batch(polys)
    ints = [kzg.fr_to_ints(polys[b]) for b in range(batch)]
    cmp_g1(commits, cref.g1_mul_gen([pyref.eval_poly(p, SECRET) for p in ints]))
    b = batch - 1
    cmp_g1(proofs[b], cref.g1_mul_gen(single_proof_exponents(ints[b], SECRET, 12)))


def test_da_using_fk20_n4096(fk4096):
    """fk20_single.go:176-196 at the BASELINE size: 8192 proofs, reverse bit order (proof i opens at
    w_8192^brp(i)), and fk20_single.go:139-172 FK20SingleDAOptimized (natural order) on the zero-padded input."""
    poly = blob_polys(1, 4096, first_blob=4242)[0]
    ints = kzg.fr_to_ints(poly)
    nat = single_proof_exponents(ints, SECRET, 13)
    rev = list(nat)
    pyref.reverse_bit_order(rev)
    got = fk4096.da_using_fk20(poly)
    assert got.shape == (8192, 18)
    cmp_g1(got, cref.g1_mul_gen(rev))
    ext = np.concatenate([poly, np.zeros_like(poly)])
    cmp_g1(fk4096.fk20_single_da_optimized(ext), cref.g1_mul_gen(nat))
    # the even positions of the natural-order result are FK20Single's proofs
    cmp_g1(fk4096.fk20_single(poly), cref.g1_mul_gen(nat[0::2]))


def test_da_using_fk20_batch_n4096(fk4096):
    """DAUsingFK20 batched (lanes = polynomials) at n = 4096: 32 polynomials, two checked in full against the closed form,
    one position of every polynomial; a ragged batch of 3 takes the per-lane program path."""
    batch = 32
    polys = blob_polys(batch, 4096, first_blob=7000)
    got = fk4096.da_using_fk20_batch(polys)
    assert got.shape == (batch, 8192, 18)
    ints = [kzg.fr_to_ints(polys[b]) for b in range(batch)]
    for b in (0, 31):
        rev = single_proof_exponents(ints[b], SECRET, 13)
        pyref.reverse_bit_order(rev)
        cmp_g1(got[b], cref.g1_mul_gen(rev))
    pos = 5000
    x = pow(pyref.scale2_root_of_unity(13), pyref.reverse_bits_limited(8192, pos), R)
    want = [(pyref.eval_poly(p, SECRET) - pyref.eval_poly(p, x)) * pow((SECRET - x) % R, -1, R) % R for p in ints]
    cmp_g1(got[:, pos], cref.g1_mul_gen(want))
    small = fk4096.da_using_fk20_batch(polys[:3])
    assert np.array_equal(kzg.g1_to_compressed(small[2]), kzg.g1_to_compressed(got[2]))


def _random_points(n, seed, affine):
    rng = random.Random(seed)
    ks = [rng.randrange(R) for _ in range(n)]
    pts = cref.g1_mul_gen(ks)
    if affine:
        pts = cref.g1_decompress(cref.g1_compress(pts))          # Z = 1
    return ks, pts


@pytest.mark.parametrize("n,affine", [(32, True), (33, False), (257, False), (600, True), (1024, False), (4096, True), (4096, False), (8192, True), (65536, True)])
def test_lincomb_bucket_msm(n, affine):
    """generic b200_g1_lincomb (no settings, no fixed-base table) over arbitrary points: the Pippenger path from 32
    terms on; affine inputs take mixed additions, Jacobian inputs (Z != 1) general ones.  Structured cases: zero
    and r - 1 scalars, infinity, the same point twice, P and -P."""
    rng = random.Random(n * 2 + affine)
    ks, pts = _random_points(n, n, affine)
    ss = [rng.randrange(R) for _ in range(n)]
    ss[1], ss[3] = 0, R - 1
    pts[2] = 0; ks[2] = 0                                        # infinity
    pts[5] = pts[4]; ks[5] = ks[4]; ss[5] = ss[4]                # equal terms meet in every bucket
    pts[7] = pts[6]; ks[7] = ks[6]; ss[7] = (R - ss[6]) % R      # k P + (r - k) P
    out = kzg.lincomb_g1(pts, kzg.fr_from_ints(ss))
    cmp_g1(out.reshape(1, 18), cref.g1_mul_gen([sum(k * s for k, s in zip(ks, ss)) % R]))
    if n <= 4096:
        cmp_g1(out.reshape(1, 18), cref.lincomb_g1(pts, cref.fr_to_limbs(ss)).reshape(1, 18))


def test_commit_without_table_uses_the_bucket_msm(trusted_setup_bytes):
    """CommitToPoly of ONE polynomial is below the fixed-base table threshold only for n < 4096; the multi-GPU
    partial commitment (b200_commit_partial_dev) always takes the generic path: point ranges of the trusted
    setup against the closed form."""
    import torch
    s1, _ = trusted_setup_bytes
    fs = kzg.FFTSettings(12)
    ks = kzg.KZGSettings(fs, kzg.g1_from_compressed(s1))
    coeffs = random_fr_ints(4096, 0xC0FFEE)
    d = torch.from_numpy(kzg.fr_from_ints(coeffs).view(np.int64)).cuda()
    L = kzg.lib()
    for lo, hi in ((0, 4096), (100, 1000), (4000, 4031)):
        out = torch.zeros((1, 18), dtype=torch.int64, device="cuda")
        assert L.b200_commit_partial_dev(ks.h, d.data_ptr(), lo, hi, out.data_ptr(), None) == 0
        torch.cuda.synchronize()
        want = sum(coeffs[i] * pow(SECRET, i, R) for i in range(lo, hi)) % R
        cmp_g1(out.cpu().numpy().view(np.uint64), cref.g1_mul_gen([want]))
    short = ks.commit_to_poly(kzg.fr_from_ints(coeffs[:1000]))
    cmp_g1(short.reshape(1, 18), cref.g1_mul_gen([pyref.eval_poly(coeffs[:1000], SECRET)]))


@pytest.mark.parametrize("scale", [12, 13])
def test_fft_g1_forward_and_inverse_at_full_size(scale):
    """fft_g1.go:58-94 at 4096 and 8192 points, both directions, in the exponent: FFTG1([k_i G]) == [FFT(k)_i G]."""
    n = 1 << scale
    ks, pts = _random_points(n, 1000 + scale, False)
    pts[3] = 0; ks[3] = 0
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    for inv in (False, True):
        want = cref.limbs_to_fr(fo.fft(cref.fr_to_limbs(ks), inv))
        cmp_g1(fs.fft_g1(pts, inv), cref.g1_mul_gen(want))


def test_config4_das_extension_and_recovery_batched():
    """config 4: DASFFTExtension + RecoverPolyFromSamples at n = 2^14, half of the samples missing, 8 polynomials
    per call.  Extension against the oracle; recovery must return exactly the extended data (round trip)."""
    scale, batch = 14, 8
    n = 1 << scale
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    even = np.stack([kzg.fr_from_ints(random_fr_ints(n // 2, 0xD4000000 + b)) for b in range(batch)])
    odd = fs.das_fft_extension_batch(even)
    for b in (0, 3, 7):
        assert np.array_equal(odd[b], fo.das_fft_extension(even[b]))
    full = np.empty((batch, n, 4), dtype=np.uint64)
    full[:, 0::2], full[:, 1::2] = even, odd
    present = np.ones((batch, n), dtype=np.uint8)
    for b in range(batch):
        rng = random.Random(14 + b)                              # recover_from_samples_bench_test.go:49 seeds with the scale
        perm = list(range(n))
        rng.shuffle(perm)
        present[b, perm[: n // 2]] = 0
    samples = full.copy()
    samples[present == 0] = 0
    rec = fs.recover_poly_from_samples_batch(samples, present)
    assert np.array_equal(rec, full)
    ze, zp = fs.zero_poly_via_multiplication(np.flatnonzero(present[2] == 0), n)
    ze_o, zp_o = fo.zero_poly(np.flatnonzero(present[2] == 0), n)
    assert np.array_equal(ze, ze_o) and np.array_equal(zp, zp_o)


def test_inplace_fft_error_split():
    """fft_fr.go:76-105 InplaceFFT: same transform as FFT for power-of-two lengths, `error` for other lengths
    (:81-83) and for too many values (:78-80); FFT pads instead (fft_fr.go:60)."""
    fs, fo = kzg.FFTSettings(6), cref.FFTSettings(6)
    v = kzg.fr_from_ints(random_fr_ints(32, 5))
    for inv in (False, True):
        assert np.array_equal(fs.inplace_fft(v, inv), fo.fft(v, inv))
    with pytest.raises(kzg.KZGError) as e:
        fs.inplace_fft(v[:24])
    assert e.value.status == kzg.kzg.NOT_POW2
    assert fs.fft(v[:24]).shape == (32, 4)
    with pytest.raises(kzg.KZGError):
        fs.inplace_fft(kzg.fr_from_ints(list(range(128))))
    with pytest.raises(kzg.KZGPanic):                     # n == 0: the reference divides by n
        fs.inplace_fft(np.zeros((0, 4), dtype=np.uint64))
    with pytest.raises(kzg.KZGPanic):                     # same in FFTG1 (fft_g1.go:78,88)
        fs.fft_g1(np.zeros((0, 18), dtype=np.uint64))


# ------------------------------------------------------------------------------ SURVEY.md 8f ranks 2-4
def test_from_compressed_on_device(trusted_setup_bytes):
    """bls/bls_kilic.go:118-121 over arrays on the GPU == the host level-1 decoder == the oracle: the 4096 + 4096
    points of eth/trusted_setup.json, infinity, and the rejections (bad flags, x >= p, not on the curve, on the curve
    but outside the prime-order subgroup)."""
    s1, lag = trusted_setup_bytes
    both = np.concatenate([s1, lag])
    got = kzg.g1_from_compressed_device(both)
    assert np.array_equal(got, cref.g1_decompress(both))
    assert np.array_equal(got[:64], kzg.g1_from_compressed(both[:64]))
    assert (got[:, 12] == 1).all() and not got[:, 13:].any()            # Z = 1
    P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    inf = np.zeros(48, dtype=np.uint8); inf[0] = 0xC0
    cases, want_ok = [s1[1], inf], [True, True]
    bad_inf = inf.copy(); bad_inf[47] = 1
    no_flag = s1[1].copy(); no_flag[0] &= 0x7F
    too_big = np.frombuffer(bytes([0x80 | (P_MOD >> 376)]) + (P_MOD & ((1 << 376) - 1)).to_bytes(47, "big"), dtype=np.uint8)
    cases += [bad_inf, no_flag, too_big]; want_ok += [False, False, False]
    n_off_curve = n_off_group = 0
    for x in range(1, 30):
        enc = np.frombuffer(bytes([0x80]) + x.to_bytes(48, "big")[1:], dtype=np.uint8)
        on_curve = pow((x ** 3 + 4) % P_MOD, (P_MOD - 1) // 2, P_MOD) == 1
        cases.append(enc); want_ok.append(False)                         # small x: never in the subgroup
        n_off_curve += not on_curve; n_off_group += on_curve
    assert n_off_curve >= 5 and n_off_group >= 5
    pts, ok = kzg.g1_from_compressed_device(np.stack(cases), return_ok=True)
    assert list(ok) == want_ok
    assert not pts[2:].any() and not pts[1].any()
    for enc, good in zip(cases, want_ok):                                # the host decoder and the oracle agree case by case
        rc = kzg.lib().b200_g1_from_compressed(np.zeros(18, dtype=np.uint64).ctypes.data, np.ascontiguousarray(enc).ctypes.data)
        assert (rc == 0) == good
    with pytest.raises(kzg.KZGError):
        kzg.g1_from_compressed_device(np.stack(cases))


def test_trusted_setup_loader_and_text_formats(trusted_setup_bytes):
    """eth/globals.go:33-49 on a trusted_setup.json-shaped document (built from the reference's own fixture bytes):
    hex UnmarshalText (bls/bls_all.go:25-39) -> device FromCompressedG1, Lagrange half in reverse bit order (:48);
    MarshalText round trip; the loaded halves satisfy the file's relation IFFT_G1(setup_G1) == setup_G1_lagrange."""
    import json
    s1, lag = trusted_setup_bytes
    doc = json.dumps({"setup_G1": [bytes(r).hex() for r in s1], "setup_G2": ["c0" + "00" * 95],
                      "setup_G1_lagrange": [bytes(r).hex() for r in lag]})
    ts = kzg.load_trusted_setup(doc)
    assert ts["setup_G2"].shape == (1, 96)
    assert np.array_equal(kzg.g1_to_compressed(ts["setup_G1"]), s1)
    brp = kzg.bit_reversal_permutation(np.arange(4096))
    assert np.array_equal(kzg.g1_to_compressed(ts["setup_G1_lagrange"]), lag[brp])
    assert kzg.g1_marshal_text(ts["setup_G1"][:5]) == [bytes(r).hex() for r in s1[:5]]
    fs = kzg.FFTSettings(12)
    nat = fs.fft_g1(ts["setup_G1"], True)
    assert np.array_equal(kzg.g1_to_compressed(nat[brp]), kzg.g1_to_compressed(ts["setup_G1_lagrange"]))
    with pytest.raises(kzg.KZGError):
        kzg.g1_unmarshal_text(["zz" * 48])
    with pytest.raises(kzg.KZGError):
        kzg.g1_unmarshal_text(["00" * 48])                               # compression flag missing
    with pytest.raises(kzg.KZGError):
        kzg.g1_unmarshal_text(["c0" + "00" * 46])                        # 47 bytes


@pytest.mark.parametrize("n,rev", [(4, False), (16, True), (4096, True), (4096, False), (512, False)])
def test_evaluate_poly_in_evaluation_form(n, rev):
    """bls/globals.go:106-153 stand-alone: y == p(x) for the interpolating polynomial (oracle inverse FFT + Horner,
    independent of the barycentric formula), natural and reverse-bit-order domains, sub-size domains (scale > 0),
    and x inside the domain (the reference's formula then yields 0)."""
    bits = n.bit_length() - 1
    scale = max(bits, 4) + (1 if n == 512 else 0)                        # n = 512 on a 1024-domain: roots stride 2
    fs, ofs = kzg.FFTSettings(scale), cref.FFTSettings(bits)
    batch = 3
    polys = np.zeros((batch, n, 4), dtype=np.uint64)
    xs, want = [], []
    w = pyref.scale2_root_of_unity(bits)
    perm = [pyref.reverse_bits_limited(n, i) for i in range(n)] if rev else list(range(n))
    for b in range(batch):
        v = random_fr_ints(n, 0xEF000000 + 16 * n + b)
        coeffs = cref.limbs_to_fr(ofs.fft(cref.fr_to_limbs(v), True))
        x = random_fr_ints(1, 0xEE000000 + b)[0] if b != 1 else pow(w, 3 % n, R)
        xs.append(x)
        want.append(pyref.eval_poly(coeffs, x) if b != 1 else 0)
        polys[b] = kzg.fr_from_ints([v[perm[i]] for i in range(n)])
    y = fs.evaluate_poly_in_evaluation_form(polys, kzg.fr_from_ints(xs), rev)
    assert kzg.fr_to_ints(y) == want


def test_check_proof_single_and_multi_g1_side(goldens):
    """kzg_single_proofs.go:57-75 and kzg_multi_proofs.go:47-88 without the pairing: the G1 operands the reference
    hands to PairingsVerify.  With the secret known, e(A, [1]) == e(proof, [t]) is A == t * proof in G1; the proofs
    come from the closed forms the reference's tests check (kzg_single_proofs_test.go / kzg_multi_proofs_test.go:
    quotient by X - x resp. X^n - x^n evaluated at s)."""
    secret = int(goldens["fk20_single_test"]["secret"])
    scale, width = 6, 64
    fs = kzg.FFTSettings(scale)
    setup = cref.generate_setup_g1(secret, width + 1)
    ks = kzg.KZGSettings(fs, setup)
    rng = random.Random(99)
    poly = [rng.randrange(R) for _ in range(32)]
    commit = ks.commit_to_poly(kzg.fr_from_ints(poly))
    ps = pyref.eval_poly(poly, secret)
    L = kzg.lib()
    # --- single: x arbitrary, y = p(x), proof = (p(s) - y) / (s - x) G
    xs = [rng.randrange(R) for _ in range(5)]
    ys = [pyref.eval_poly(poly, x) for x in xs]
    a = kzg.check_proof_single_g1(np.stack([commit] * 5), kzg.fr_from_ints(ys))
    cmp_g1(a, cref.g1_mul_gen([(ps - y) % R for y in ys]))
    proofs = cref.g1_mul_gen([(ps - y) * pow((secret - x) % R, -1, R) % R for x, y in zip(xs, ys)])
    for i in range(5):                                                    # the pairing equation, in G1
        t = kzg.fr_from_ints([(secret - xs[i]) % R])
        lhs = np.zeros(18, dtype=np.uint64)
        L.b200_g1_mul(lhs.ctypes.data, proofs[i].ctypes.data, t.ctypes.data)
        assert L.b200_g1_equal(lhs.ctypes.data, np.ascontiguousarray(a[i]).ctypes.data) == 1
    wrong = kzg.check_proof_single_g1(commit.reshape(1, 18), kzg.fr_from_ints([(ys[0] + 1) % R]))
    assert L.b200_g1_equal(np.ascontiguousarray(wrong[0]).ctypes.data, np.ascontiguousarray(a[0]).ctypes.data) == 0
    # --- multi: coset x <w_n>, n = 8: ys[i] = p(x w^i)
    n = 8
    wn = pyref.scale2_root_of_unity(3)
    ofs = pyref.FFTSettings(3)
    batch = 4
    xs = [rng.randrange(1, R) for _ in range(batch)]
    xs[1] = 0                                                             # InvModFr(0) = 0: only the constant coefficient survives
    ys_all, want_is, want_xn = [], [], []
    for x in xs:
        ysb = [pyref.eval_poly(poly, x * pow(wn, i, R) % R) for i in range(n)]
        ys_all.append(kzg.fr_from_ints(ysb))
        interp = ofs.fft(ysb, True)
        xinv = pow(x, -1, R) if x else 0
        interp = [c * pow(xinv, i, R) % R for i, c in enumerate(interp)]
        want_is.append(pyref.eval_poly(interp, secret))
        want_xn.append(pow(x, n, R))
    got, xn = ks.check_proof_multi_g1(np.stack([commit] * batch), kzg.fr_from_ints(xs), np.stack(ys_all))
    assert kzg.fr_to_ints(xn) == want_xn
    cmp_g1(got, cref.g1_mul_gen([(ps - v) % R for v in want_is]))
    for b in (0, 2, 3):                                                   # p - I == q (X^n - x^n): A == (s^n - x^n) * [q(s)]
        q_at_s = (ps - want_is[b]) * pow((pow(secret, n, R) - want_xn[b]) % R, -1, R) % R
        proof = cref.g1_mul_gen([q_at_s])[0]
        t = kzg.fr_from_ints([(pow(secret, n, R) - want_xn[b]) % R])
        lhs = np.zeros(18, dtype=np.uint64)
        L.b200_g1_mul(lhs.ctypes.data, proof.ctypes.data, t.ctypes.data)
        assert L.b200_g1_equal(lhs.ctypes.data, np.ascontiguousarray(got[b]).ctypes.data) == 1
    # a large batch takes the fixed-base table route (batch * n >= 4096): same values
    big = 600
    xs2 = [rng.randrange(1, R) for _ in range(big)]
    ys2 = np.stack([kzg.fr_from_ints([pyref.eval_poly(poly[:4], x * pow(wn, i, R) % R) for i in range(n)]) for x in xs2])
    commit4 = ks.commit_to_poly(kzg.fr_from_ints(poly[:4]))
    got2, _ = ks.check_proof_multi_g1(np.stack([commit4] * big), kzg.fr_from_ints(xs2), ys2)
    cmp_g1(got2, np.zeros((big, 18), dtype=np.uint64))                    # deg p < n: the interpolation IS p, difference = infinity
    with pytest.raises(kzg.KZGPanic):
        ks.check_proof_multi_g1(commit.reshape(1, 18), kzg.fr_from_ints([5]), np.zeros((1, 128, 4), dtype=np.uint64))


def test_toeplitz_part2_part3_standalone(goldens):
    """fk20_single.go:59-87 as stand-alone methods == the oracle's FK20Single internals: Part3(Part2(coeffs, xExtFFT))
    is h, and FFTG1(h) the proofs (fk20_single.go:122-134)."""
    t = goldens["fk20_single_test"]
    secret, poly, scale, n2 = int(t["secret"]), t["poly"], t["fft_scale"], t["n2"]
    n = len(poly)
    setup = cref.generate_setup_g1(secret, n2 + 1)
    fs = kzg.FFTSettings(scale)
    fk = kzg.FK20SingleSettings(kzg.KZGSettings(fs, setup), n2)
    coeffs = kzg.fr_from_ints(pyref.toeplitz_coeffs_step(poly))
    h_ext = fs.toeplitz_part2(coeffs, fk.x_ext_fft())
    ofs = pyref.FFTSettings(scale)
    x = [pow(secret, j, R) for j in range(n - 2, -1, -1)] + [0] * (n + 1)
    want_ext = [a * b % R for a, b in zip(ofs.fft(pyref.toeplitz_coeffs_step(poly)), ofs.fft(x))]
    cmp_g1(h_ext, cref.g1_mul_gen(want_ext))
    h = fs.toeplitz_part3(h_ext)
    assert h.shape == (n, 18)
    cmp_g1(h, cref.g1_mul_gen(ofs.fft(want_ext, True)[:n]))
    cmp_g1(fs.fft_g1(h), fk.fk20_single(kzg.fr_from_ints(poly)))
    with pytest.raises(kzg.KZGPanic):                                     # fk20_single.go:60-62
        fs.toeplitz_part2(coeffs[:16], fk.x_ext_fft())


def test_fk20_multi_settings_sharded_by_offset():
    """Per-rank FK20 multi settings (only the files of the rank's chunk offsets): partial hExtFFT per rank -> G1 sum
    -> finish == DAUsingFK20Multi of full settings; whole-polynomial entry points refuse a partial handle."""
    import torch
    from go_kzg_b200 import multi_gpu
    secret, l, cc = 1927409816240961209460912649124, 16, 32
    n = l * cc
    scale = (2 * n).bit_length() - 1
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, kzg.generate_testing_setup_g1(secret, 2 * n))
    full = kzg.FK20MultiSettings(ks, 2 * n, l)
    poly = kzg.fr_from_ints(random_fr_ints(n, 31))
    want = full.da_using_fk20_multi(poly)
    L, k2, world = kzg.lib(), 2 * cc, 3
    d_poly = torch.from_numpy(poly.view(np.int64)).cuda()
    parts = torch.zeros((world, k2, 18), dtype=torch.int64, device="cuda")
    last = None
    for r in range(world):
        mine = multi_gpu.offset_range(r, world, l)
        fk_r = kzg.FK20MultiSettings(ks, 2 * n, l, offsets=mine)
        cmp_g1(fk_r.x_ext_fft(mine.start), full.x_ext_fft(mine.start))
        assert L.b200_fk20_multi_partial_dev(fk_r.h, d_poly.data_ptr(), n, mine.start, mine.stop, parts[r].data_ptr(), None) == 0
        other = (mine.stop % l)
        if other not in mine:
            assert L.b200_fk20_multi_partial_dev(fk_r.h, d_poly.data_ptr(), n, other, other + 1, parts[r].data_ptr(), None) != 0
        last = fk_r
    d_sum = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    assert L.b200_g1_sum_dev(parts.data_ptr(), world, k2, d_sum.data_ptr(), None) == 0
    d_out = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    assert L.b200_fk20_multi_finish_dev(last.h, d_sum.data_ptr(), 1, d_out.data_ptr(), None) == 0
    torch.cuda.synchronize()
    cmp_g1(d_out.cpu().numpy().view(np.uint64), want)
    with pytest.raises(kzg.KZGPanic):
        last.da_using_fk20_multi(poly)


@pytest.mark.parametrize("latency_mode", [1, 0])
def test_single_transform_lane_mappings_agree(latency_mode):
    """One transform uses several lane mappings over its stages: a quad of lanes per butterfly while the call has the
    device to itself (latency mode 1, small launches), else one lane per butterfly -- across blocks with sparse programs
    while a stage has enough blocks, per-lane fixed windows for the rest; batches of 16+ use whole-warp lanes.  Same input
    through batch 1, 3 (below the whole-warp switch) and 16, with and without the quad kernels, must give identical bytes,
    at a size with many across-block stages."""
    n, scale = 2048, 11
    ks, pts = _random_points(n, 2048, False)
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    assert kzg.lib().b200_set_latency_mode(latency_mode) == 0
    try:
        for inv in (False, True):
            want = cref.g1_compress(cref.g1_mul_gen(cref.limbs_to_fr(fo.fft(cref.fr_to_limbs(ks), inv))))
            assert np.array_equal(kzg.g1_to_compressed(fs.fft_g1(pts, inv)), want)
            for batch in (3, 16):
                out = fs.fft_g1_batch(np.stack([pts] * batch), inv)
                assert np.array_equal(kzg.g1_to_compressed(out[batch - 1]), want)
    finally:
        kzg.lib().b200_set_latency_mode(1)
    assert kzg.lib().b200_set_latency_mode(7) != 0


def test_one_polynomial_calls_same_bytes_in_both_latency_modes(fk4096):
    """FK20Single / DAUsingFK20 of ONE polynomial at n = 4096: the quad-per-butterfly stage kernels (latency mode 1, default)
    and the one-lane kernels (mode 0, also what concurrent callers get) must produce identical proofs; spot-checked against
    the closed form."""
    rng = random.Random(31337)
    poly = [rng.randrange(R) for _ in range(4096)]
    p = kzg.fr_from_ints(poly)
    L = kzg.lib()
    out = {}
    try:
        for mode in (1, 0):
            assert L.b200_set_latency_mode(mode) == 0
            out[mode] = (kzg.g1_to_compressed(fk4096.fk20_single(p)), kzg.g1_to_compressed(fk4096.da_using_fk20(p)))
    finally:
        L.b200_set_latency_mode(1)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    ps = pyref.eval_poly(poly, SECRET)
    w = pyref.scale2_root_of_unity(12)
    for i in (0, 1, 2047, 4095):
        x = pow(w, i, R)
        q = (ps - pyref.eval_poly(poly, x)) * pow((SECRET - x) % R, -1, R) % R
        assert bytes(out[1][0][i]) == bytes(cref.g1_compress(cref.g1_mul_gen([q]))[0])


@pytest.mark.parametrize("scale", [2, 5, 8, 12])
def test_das_fft_extension_over_g1(scale):
    """The G1 form of DASFFTExtension (das_extension.go:7-84; the TODO at fk20_multi.go:96), in the exponent:
    ext([k_i G]) == [ext(k)_i G] with the oracle's Fr extension; interleaved, the 2n points are the FFTG1 of a
    coefficient vector whose upper half is infinity (das_extension_test.go:42-86 over G1)."""
    n = 1 << (scale - 1)
    ks, pts = _random_points(n, 77 + scale, scale % 2 == 0)
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    odd = fs.das_fft_extension_g1(pts)
    want = cref.limbs_to_fr(fo.das_fft_extension(cref.fr_to_limbs(ks)))
    cmp_g1(odd, cref.g1_mul_gen(want))
    if scale <= 8:
        inter = np.empty((2 * n, 18), dtype=np.uint64)
        inter[0::2], inter[1::2] = pts, odd
        coeffs = fs.fft_g1(inter, True)
        assert (kzg.g1_to_compressed(coeffs[n:])[:, 0] == 0xC0).all()
    with pytest.raises(kzg.KZGPanic):
        kzg.FFTSettings(3).das_fft_extension_g1(np.zeros((8, 18), dtype=np.uint64))


def test_fk20_settings_disk_cache(tmp_path, goldens):
    """xExtFFT as a cached artefact.  The second construction adopts the stored files
    (b200_fk20_settings_new_from_x_ext_fft) and proves identically; another setup gets another key."""
    t = goldens["fk20_multi_test"]
    secret, l, cc = int(t["secret"]), t["chunk_len"], t["chunk_count"]
    n = l * cc
    scale = (2 * n).bit_length() - 1
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, cref.generate_setup_g1(secret, 2 * n))
    poly = kzg.fr_from_ints(random_fr_ints(n, 41))
    a = kzg.FK20MultiSettings(ks, 2 * n, l, cache_dir=str(tmp_path))
    b = kzg.FK20MultiSettings(ks, 2 * n, l, cache_dir=str(tmp_path))
    assert not a.from_cache and b.from_cache
    want = a.da_using_fk20_multi(poly)
    assert np.array_equal(kzg.g1_to_compressed(b.da_using_fk20_multi(poly)), kzg.g1_to_compressed(want))
    cmp_g1(b.x_ext_fft(5), a.x_ext_fft(5))
    s1 = kzg.FK20SingleSettings(ks, 2 * n, cache_dir=str(tmp_path))
    s2 = kzg.FK20SingleSettings(ks, 2 * n, cache_dir=str(tmp_path))
    assert not s1.from_cache and s2.from_cache
    cmp_g1(s2.fk20_single(poly), cref.g1_mul_gen(pyref.fk20_single_exponents(kzg.fr_to_ints(poly), secret)))
    ks2 = kzg.KZGSettings(fs, cref.generate_setup_g1(secret + 1, 2 * n))
    c = kzg.FK20MultiSettings(ks2, 2 * n, l, cache_dir=str(tmp_path))
    assert not c.from_cache
    r = kzg.FK20MultiSettings(ks, 2 * n, l, offsets=range(4, 9), cache_dir=str(tmp_path))
    r2 = kzg.FK20MultiSettings(ks, 2 * n, l, offsets=range(4, 9), cache_dir=str(tmp_path))
    assert not r.from_cache and r2.from_cache
    cmp_g1(r2.x_ext_fft(8), a.x_ext_fft(8))


def test_concurrent_callers_share_handles(fk4096):
    """include/b200_kzg.h: handles may be used from any thread concurrently (goroutines in the Go binding).  Eight host
    threads hammer one FK20 settings object, one KZGSettings (lazily built commitment table) and the handle-less
    LinCombG1 at once -- different entry points, different sizes -- and every result must equal the sequential one."""
    import threading
    polys = blob_polys(8, 4096, first_blob=12000)
    ks_pts, pts = _random_points(300, 4321, True)
    scal = kzg.fr_from_ints([random.Random(5).randrange(R) for _ in range(300)])
    want_single = [kzg.g1_to_compressed(fk4096.fk20_single(polys[i])) for i in range(4)]
    want_commit = kzg.g1_to_compressed(fk4096.ks.commit_to_poly_batch(polys))
    want_lin = kzg.g1_to_compressed(kzg.lincomb_g1(pts, scal).reshape(1, 18))
    want_fft = fk4096.ks.fs.fft(polys[0])
    errors = []

    def worker(t):
        try:
            for rep in range(3):
                if t < 4:
                    got = kzg.g1_to_compressed(fk4096.fk20_single(polys[t]))
                    assert np.array_equal(got, want_single[t]), "fk20_single thread %d" % t
                elif t < 6:
                    got = kzg.g1_to_compressed(fk4096.ks.commit_to_poly_batch(polys))
                    assert np.array_equal(got, want_commit), "commit thread %d" % t
                elif t == 6:
                    got = kzg.g1_to_compressed(kzg.lincomb_g1(pts, scal).reshape(1, 18))
                    assert np.array_equal(got, want_lin), "lincomb"
                else:
                    assert np.array_equal(fk4096.ks.fs.fft(polys[0]), want_fft), "fft"
        except Exception as e:      # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("n,batch", [(1, 9), (2, 4), (16, 5), (64, 70), (256, 33), (1024, 7), (1000, 6)])
def test_fft_fr_batches_of_short_transforms(n, batch):
    """Short transforms are packed several to a CTA tile (kernels_fr.cu: launch_fr_ntt): full groups plus a ragged tail,
    both directions, against the oracle; n = 1000 pads to 1024 like fft_fr.go:60."""
    scale = max((n - 1).bit_length(), 4)
    fs, fo = kzg.FFTSettings(scale), cref.FFTSettings(scale)
    v = np.stack([kzg.fr_from_ints(random_fr_ints(n, 31 * n + b)) for b in range(batch)])
    for inv in (False, True):
        out = fs.fft_batch(v, inv)
        for b in (0, batch // 2, batch - 1):
            assert np.array_equal(out[b], fo.fft(v[b], inv))
