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


def _random_points(n, seed, affine):
    rng = random.Random(seed)
    ks = [rng.randrange(R) for _ in range(n)]
    pts = cref.g1_mul_gen(ks)
    if affine:
        pts = cref.g1_decompress(cref.g1_compress(pts))          # Z = 1
    return ks, pts


@pytest.mark.parametrize("n,affine", [(32, True), (33, False), (257, False), (4096, True), (4096, False), (65536, True)])
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
