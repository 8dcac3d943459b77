"""Prove -> verify through the pairing, the way the reference's own tests close the loop (fk20_single_test.go:11-45,
fk20_multi_test.go:11-103, kzg_single_proofs_test.go / kzg_multi_proofs_test.go): proofs from the device FK20 paths are
checked with KZGSettings.CheckProofSingle / CheckProofMulti (G1 sides on the device, G2 arithmetic and pairing in the
library's host code) -- no known-secret shortcut on the verification side."""
import random

import numpy as np
import pytest

import go_kzg_b200 as kzg
from oracle import cref, pyref

pytestmark = pytest.mark.gpu
R = pyref.R_MOD


def rev_bits(n, i):
    return int(bin(i)[2:].zfill(n.bit_length() - 1)[::-1], 2) if n > 1 else 0


def test_da_using_fk20_verifies_like_the_reference_test(goldens):
    """fk20_single_test.go:11-45 TestKZGSettings_DAUsingFK20, every position instead of only pos = 9"""
    secret = int(goldens["fk20_single_test"]["secret"])
    poly = goldens["fk20_single_test"]["poly"]
    fs = kzg.FFTSettings(5)
    s1, s2 = cref.generate_setup_g1(secret, 33), kzg.generate_testing_setup_g2(secret, 33)
    ks = kzg.KZGSettings(fs, s1, secret_g2=s2)
    fk = kzg.FK20SingleSettings(ks, 32)
    commitment = ks.commit_to_poly(kzg.fr_from_ints(poly))
    proofs = fk.da_using_fk20(kzg.fr_from_ints(poly))
    roots = kzg.fr_to_ints(fs.expanded_roots_of_unity())
    for pos in range(32):
        x = roots[pos]
        y = pyref.eval_poly(poly, x)
        proof = proofs[rev_bits(32, pos)]
        assert ks.check_proof_single(commitment, proof, kzg.fr_from_ints([x]), kzg.fr_from_ints([y]))
    x, y = roots[9], pyref.eval_poly(poly, roots[9])
    assert not ks.check_proof_single(commitment, proofs[rev_bits(32, 9)], kzg.fr_from_ints([x]), kzg.fr_from_ints([(y + 1) % R]))
    assert not ks.check_proof_single(commitment, proofs[rev_bits(32, 10)], kzg.fr_from_ints([x]), kzg.fr_from_ints([y]))
    # batch form: all 32 at once, one of them with the wrong y
    xs = [roots[p] for p in range(32)]
    ys = [pyref.eval_poly(poly, v) for v in xs]
    ys[5] = (ys[5] + 7) % R
    ok = ks.check_proof_single_batch(np.stack([commitment] * 32), np.stack([proofs[rev_bits(32, p)] for p in range(32)]),
                                     kzg.fr_from_ints(xs), kzg.fr_from_ints(ys))
    assert list(ok) == [p != 5 for p in range(32)]
    # aggregated form: one pairing for the batch
    good = [pyref.eval_poly(poly, v) for v in xs]
    cs, ps = np.stack([commitment] * 32), np.stack([proofs[rev_bits(32, p)] for p in range(32)])
    assert ks.check_proof_single_aggregate(cs, ps, kzg.fr_from_ints(xs), kzg.fr_from_ints(good))
    assert not ks.check_proof_single_aggregate(cs, ps, kzg.fr_from_ints(xs), kzg.fr_from_ints(ys))
    assert ks.check_proof_single_aggregate(cs, ps, kzg.fr_from_ints(xs), kzg.fr_from_ints(good), rs=kzg.fr_from_ints(range(1, 33)))
    with pytest.raises(kzg.KZGPanic):                               # a zero coefficient would drop a proof from the check
        ks.check_proof_single_aggregate(cs, ps, kzg.fr_from_ints(xs), kzg.fr_from_ints(good), rs=kzg.fr_from_ints(range(32)))
    # without SecretG2 the check cannot run
    bare = kzg.KZGSettings(fs, s1)
    with pytest.raises(kzg.KZGPanic):
        bare.check_proof_single(commitment, proofs[0], kzg.fr_from_ints([x]), kzg.fr_from_ints([y]))


def test_da_using_fk20_multi_verifies_like_the_reference_test(goldens):
    """fk20_multi_test.go:11-103 TestKZGSettings_DAUsingFK20Multi: every coset proof through CheckProofMulti"""
    g = goldens["fk20_multi_test"]
    secret, l, cc = int(g["secret"]), g["chunk_len"], g["chunk_count"]
    n = l * cc
    fs = kzg.FFTSettings(g["fft_scale"])
    s1 = cref.generate_setup_g1(secret, 2 * n)
    s2 = kzg.generate_testing_setup_g2(secret, l + 1)             # CheckProofMulti reads SecretG2[len(ys)] only
    ks = kzg.KZGSettings(fs, s1, secret_g2=s2)
    fk = kzg.FK20MultiSettings(ks, 2 * n, l)
    poly = []
    for i in range(cc):
        row = [1, 2, 3, 4 + i, 7, 8 + i * i, 9, 10, 13, 14, 1, 15, R - 1, 1000, R - 134, 33]
        poly += row
    commitment = ks.commit_to_poly(kzg.fr_from_ints(poly))
    proofs = fk.da_using_fk20_multi(kzg.fr_from_ints(poly))
    assert proofs.shape[0] == 2 * cc
    ext = kzg.fr_to_ints(fs.fft(kzg.fr_from_ints(poly + [0] * n)))
    ext = [ext[rev_bits(2 * n, i)] for i in range(2 * n)]           # reverseBitOrderFr(extendedData)
    roots = kzg.fr_to_ints(fs.expanded_roots_of_unity())
    stride = (1 << g["fft_scale"]) // (2 * n)
    xs, ys_all = [], []
    for pos in range(2 * cc):
        x = roots[rev_bits(2 * cc, pos) * stride]
        ys = ext[l * pos:l * (pos + 1)]
        ys = [ys[rev_bits(l, i)] for i in range(l)]                 # reverseBitOrderFr(ys)
        if pos in (0, 17):                                          # ... and they are the evaluations on the coset
            w = roots[(1 << g["fft_scale"]) // l]
            assert ys == [pyref.eval_poly(poly, x * pow(w, i, R) % R) for i in range(l)]
        xs.append(x)
        ys_all.append(kzg.fr_from_ints(ys))
    for pos in (0, 1, 31, 63):
        assert ks.check_proof_multi(commitment, proofs[pos], kzg.fr_from_ints([xs[pos]]), ys_all[pos])
    bad = ys_all[3].copy()
    bad[2, 0] ^= 1
    assert not ks.check_proof_multi(commitment, proofs[3], kzg.fr_from_ints([xs[3]]), bad)
    ok = ks.check_proof_multi_batch(np.stack([commitment] * (2 * cc)), proofs, kzg.fr_from_ints(xs), np.stack(ys_all))
    assert ok.all()
    swapped = proofs.copy()
    swapped[[10, 11]] = swapped[[11, 10]]
    ok = ks.check_proof_multi_batch(np.stack([commitment] * (2 * cc)), swapped, kzg.fr_from_ints(xs), np.stack(ys_all))
    assert list(ok) == [p not in (10, 11) for p in range(2 * cc)]
    assert ks.check_proof_multi_aggregate(np.stack([commitment] * (2 * cc)), proofs, kzg.fr_from_ints(xs), np.stack(ys_all))
    assert not ks.check_proof_multi_aggregate(np.stack([commitment] * (2 * cc)), swapped, kzg.fr_from_ints(xs), np.stack(ys_all))
    with pytest.raises(kzg.KZGPanic):                               # SecretG2[32] is not there
        ks.check_proof_multi(commitment, proofs[0], kzg.fr_from_ints([xs[0]]), np.zeros((32, 4), dtype=np.uint64))


def test_headline_proofs_verify_at_full_size(trusted_setup_bytes, goldens):
    """n = 4096 over the trusted setup (secret 1337): commitment and FK20Single proofs of a random blob, spot-checked
    with the pairing at 12 random positions; SecretG2[1] decoded from the reference's own setup_G2 entry."""
    s1, _ = trusted_setup_bytes
    first = kzg.g1_from_compressed(s1)
    rest = cref.g1_mul_gen([pow(1337, i, R) for i in range(4096, 8192)])
    g2 = kzg.g2_from_compressed(np.frombuffer(bytes.fromhex(goldens["trusted_setup_g2"]["entries"]["0"])
                                              + bytes.fromhex(goldens["trusted_setup_g2"]["entries"]["1"]), dtype=np.uint8))
    fs = kzg.FFTSettings(13)
    ks = kzg.KZGSettings(fs, np.concatenate([first, rest]), secret_g2=g2)
    fk = kzg.FK20SingleSettings(ks, 8192)
    rng = random.Random(4096)
    poly = [rng.randrange(R) for _ in range(4096)]
    commits, proofs = fk.commit_fk20_batch(kzg.fr_from_ints(poly).reshape(1, 4096, 4))
    w = pyref.scale2_root_of_unity(12)
    pos = rng.sample(range(4096), 12)
    xs = [pow(w, p, R) for p in pos]
    ys = [pyref.eval_poly(poly, x) for x in xs]
    ok = ks.check_proof_single_batch(np.stack([commits[0]] * 12), proofs[0][pos], kzg.fr_from_ints(xs), kzg.fr_from_ints(ys))
    assert ok.all()
    ok = ks.check_proof_single_batch(np.stack([commits[0]] * 12), proofs[0][[(p + 1) % 4096 for p in pos]], kzg.fr_from_ints(xs),
                                     kzg.fr_from_ints(ys))
    assert not ok.any()
    # ALL 4096 proofs with one pairing: x_i = w^i, y_i = p(w^i) = FFT(poly)[i]
    all_x = kzg.fr_from_ints([pow(w, i, R) for i in range(4096)])
    all_y = kzg.FFTSettings(12).fft(kzg.fr_from_ints(poly))
    cs = np.repeat(commits[0].reshape(1, 18), 4096, axis=0)
    assert ks.check_proof_single_aggregate(cs, proofs[0], all_x, all_y)
    tampered = proofs[0].copy()
    tampered[[100, 200]] = tampered[[200, 100]]
    assert not ks.check_proof_single_aggregate(cs, tampered, all_x, all_y)
