"""G2 arithmetic and the pairing of the library (host code, go_kzg_b200/csrc/pairing.h) -- no GPU needed.

Pinned by (a) the reference's own constants and fixtures: the G2 generator (bls/bls_hbls.go:27-30) and entries of
eth/trusted_setup.json's setup_G2 (= 1337^i * GenG2, compressed; tests/golden/reference_goldens.json), and (b) an independent
restatement of the pairing with Python integers (tests/pairing_ref.py), compared as GT elements coefficient by coefficient.
The reference's pairing engine (kilic, bls/bls_kilic.go:152-158) is not in the tree; its tests only use it through
CheckProofSingle / CheckProofMulti, which tests/test_gpu_shapes.py mirrors on the device.
"""
import numpy as np
import pytest

import go_kzg_b200 as kzg
import pairing_ref as pr

R = pr.R
P = pr.P


def g1_abi(p):
    out = np.zeros(18, dtype=np.uint64)
    if p is None:
        return out
    for c, v in enumerate((p[0], p[1], 1)):
        for j in range(6):
            out[6 * c + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def g2_abi(q):
    out = np.zeros(36, dtype=np.uint64)
    if q is None:
        return out
    for c, v in enumerate((q[0][0], q[0][1], q[1][0], q[1][1], 1, 0)):
        for j in range(6):
            out[6 * c + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def test_g2_generator_is_the_reference_constant(goldens):
    x0, x1, y0, y1 = [int(v) for v in goldens["g2_generator"]["x0x1y0y1"]]
    assert ((x0, x1), (y0, y1)) == pr.G2
    assert np.array_equal(kzg.g2_generator(), g2_abi(pr.G2))


def test_g2_compression_against_the_trusted_setup(goldens):
    """setup_G2[i] = 1337^i * GenG2 in the 96-byte ZCash form (eth/globals.go:33-49 decodes them with FromCompressedG2)"""
    entries = goldens["trusted_setup_g2"]["entries"]
    setup = kzg.generate_testing_setup_g2(1337, 4)
    for i in range(4):
        want = bytes.fromhex(entries[str(i)])
        assert bytes(kzg.g2_to_compressed(setup[i])[0]) == want
        assert pr.g2_compress(pr.g2_mul(pr.G2, pow(1337, i, R))) == want
    for i in (8, 16, 4095):
        raw = np.frombuffer(bytes.fromhex(entries[str(i)]), dtype=np.uint8)
        pt = kzg.g2_from_compressed(raw)[0]
        assert kzg.g2_equal(pt, kzg.g2_mul(kzg.g2_generator(), pow(1337, i, R)))
        assert bytes(kzg.g2_to_compressed(pt)[0]) == bytes(raw)
    inf = np.zeros(36, dtype=np.uint64)
    assert bytes(kzg.g2_to_compressed(inf)[0]) == bytes([0xC0]) + bytes(95)
    assert kzg.g2_equal(kzg.g2_from_compressed(np.frombuffer(bytes([0xC0]) + bytes(95), dtype=np.uint8))[0], inf)


def test_g2_group_law_against_python():
    g = kzg.g2_generator()
    a, b = 0x1234567890ABCDEF1234567890ABCDEF, R - 5
    qa, qb = pr.g2_mul(pr.G2, a), pr.g2_mul(pr.G2, b)
    assert kzg.g2_equal(kzg.g2_mul(g, a), g2_abi(qa))
    assert kzg.g2_equal(kzg.g2_mul(g, b), g2_abi(qb))
    assert kzg.g2_equal(kzg.g2_add(g2_abi(qa), g2_abi(qb)), g2_abi(pr.g2_add(qa, qb)))
    assert kzg.g2_equal(kzg.g2_add(g2_abi(qa), g2_abi(qa)), g2_abi(pr.g2_add(qa, qa)))          # doubling branch
    assert kzg.g2_equal(kzg.g2_sub(g2_abi(qa), g2_abi(qa)), np.zeros(36, dtype=np.uint64))      # P - P = infinity
    assert kzg.g2_equal(kzg.g2_add(g2_abi(qa), np.zeros(36, dtype=np.uint64)), g2_abi(qa))
    assert kzg.g2_equal(kzg.g2_neg(kzg.g2_neg(g2_abi(qa))), g2_abi(qa))
    assert kzg.g2_equal(kzg.g2_mul(g, 0), np.zeros(36, dtype=np.uint64))
    assert not kzg.g2_equal(g2_abi(qa), g2_abi(qb))


def _f2_sqrt(a):
    """square root in Fp2 (norm method), None if there is none -- only to build a twist point outside G2"""
    a0, a1 = a
    def fsqrt(v):
        s = pow(v, (P + 1) // 4, P)
        return s if s * s % P == v % P else None
    if a1 == 0:
        s = fsqrt(a0)
        if s is not None: return (s, 0)
        s = fsqrt(-a0 % P)
        return (0, s) if s is not None else None
    n = fsqrt((a0 * a0 + a1 * a1) % P)
    if n is None: return None
    for t in ((a0 + n) * pow(2, -1, P) % P, (a0 - n) * pow(2, -1, P) % P):
        x0 = fsqrt(t)
        if x0:
            x1 = a1 * pow(2 * x0, -1, P) % P
            if pr.f2_mul((x0, x1), (x0, x1)) == (a0 % P, a1 % P):
                return (x0, x1)
    return None


def test_g2_from_compressed_rejections(goldens):
    good = bytearray(bytes.fromhex(goldens["trusted_setup_g2"]["entries"]["1"]))
    dec = lambda b: kzg.g2_from_compressed(np.frombuffer(bytes(b), dtype=np.uint8))
    dec(good)
    bad = bytearray(good); bad[0] &= 0x7F                       # compression flag missing
    with pytest.raises(kzg.KZGError):
        dec(bad)
    bad = bytearray([0xC0] + [0] * 95); bad[50] = 1             # infinity flag with payload
    with pytest.raises(kzg.KZGError):
        dec(bad)
    bad = bytearray(good); bad[48:96] = P.to_bytes(48, "big")   # x.c0 >= p
    with pytest.raises(kzg.KZGError):
        dec(bad)
    # a point of the twist outside the order-r subgroup (the cofactor is huge: almost every curve point is one), and an x
    # with no point above it
    off_curve = on_curve_not_g2 = None
    for x0 in range(1, 200):
        x = (x0, 0)
        rhs = pr.f2_add(pr.f2_mul(pr.f2_mul(x, x), x), (4, 4))
        y = _f2_sqrt(rhs)
        if y is None:
            off_curve = off_curve or x
        elif on_curve_not_g2 is None and pr._mul(pr.g2_add, (x, y), R) is not None:
            on_curve_not_g2 = (x, y)
        if off_curve and on_curve_not_g2:
            break
    assert off_curve and on_curve_not_g2
    b = bytearray(off_curve[1].to_bytes(48, "big") + off_curve[0].to_bytes(48, "big")); b[0] |= 0x80
    with pytest.raises(kzg.KZGError):
        dec(b)
    with pytest.raises(kzg.KZGError):
        dec(pr.g2_compress(on_curve_not_g2))


def test_pairing_value_against_the_python_restatement():
    cases = [(1, 1), (0xDEADBEEFCAFEBABE0123456789, R - 2)]
    for a, b in cases:
        p, q = pr.g1_mul(pr.G1, a), pr.g2_mul(pr.G2, b)
        got = kzg.pairing(g1_abi(p), g2_abi(q))
        assert pr.gt_from_flat(got) == pr.pairing(p, q)
    one = [1] + [0] * 11
    assert kzg.pairing(g1_abi(None), g2_abi(pr.G2)) == one
    assert kzg.pairing(g1_abi(pr.G1), g2_abi(None)) == one


def test_pairings_verify_bilinearity():
    g1, g2 = g1_abi(pr.G1), kzg.g2_generator()
    a, b = 0xABCDEF0123456789ABCDEF, 0x13579BDF02468ACE
    pa = g1_abi(pr.g1_mul(pr.G1, a))
    pab = g1_abi(pr.g1_mul(pr.G1, a * b % R))
    qb = kzg.g2_mul(g2, b)
    assert kzg.pairings_verify(pa, qb, pab, g2)                 # e(aG, bH) == e(abG, H)
    assert kzg.pairings_verify(pa, qb, g1, kzg.g2_mul(g2, a * b % R))
    assert not kzg.pairings_verify(pa, qb, pab, qb)
    assert not kzg.pairings_verify(pa, g2, g1, g2)
    inf1, inf2 = np.zeros(18, dtype=np.uint64), np.zeros(36, dtype=np.uint64)
    assert kzg.pairings_verify(inf1, g2, g1, inf2)              # 1 == 1
    assert not kzg.pairings_verify(inf1, g2, g1, g2)
    off = g1.copy(); off[0] ^= 1                                # not on the curve
    with pytest.raises(kzg.KZGPanic):
        kzg.pairings_verify(off, g2, g1, g2)
    big = g1.copy(); big[0:6] = np.frombuffer(((1 << 384) - 1).to_bytes(48, "little"), dtype=np.uint64)
    with pytest.raises(kzg.KZGPanic):
        kzg.pairings_verify(big, g2, g1, g2)


def test_kzg_equation_with_the_known_secret():
    """The check of kzg_single_proofs.go:57-75 assembled from the level-1 calls, secret known:
    e([p(s) - y]_1, [1]_2) == e([(p(s) - y) / (s - x)]_1, [s - x]_2)."""
    s, x = 1927409816240961209460912649124, 0x5555
    poly = [3, 1, 4, 1, 5, 9, 2, 6]
    ev = lambda t: sum(c * pow(t, i, R) for i, c in enumerate(poly)) % R
    ps, y = ev(s), ev(x)
    lhs = g1_abi(pr.g1_mul(pr.G1, (ps - y) % R))
    proof = g1_abi(pr.g1_mul(pr.G1, (ps - y) * pow(s - x, -1, R) % R))
    g2 = kzg.g2_generator()
    s_minus_x = kzg.g2_sub(kzg.g2_mul(g2, s), kzg.g2_mul(g2, x))
    assert kzg.pairings_verify(lhs, g2, proof, s_minus_x)
    wrong = g1_abi(pr.g1_mul(pr.G1, (ps - y + 1) % R))
    assert not kzg.pairings_verify(wrong, g2, proof, s_minus_x)


def test_affine_coordinates_for_str():
    """StrG1 / StrG2 (bls/bls_kilic.go:55-61,96-102) print the affine coordinates: b200_g1_to_affine / b200_g2_to_affine"""
    L = kzg.lib()
    k = 0xFEDCBA9876543210
    p, q = pr.g1_mul(pr.G1, k), pr.g2_mul(pr.G2, k)
    # hand the library projectively scaled points: (X, Y, Z) = (x z^2, y z^3, z)
    z = 0x123456789
    pj = np.zeros(18, dtype=np.uint64)
    for c, v in enumerate((p[0] * z * z % P, p[1] * z ** 3 % P, z)):
        for j in range(6):
            pj[6 * c + j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    xy = np.zeros(12, dtype=np.uint64)
    L.b200_g1_to_affine(pj.ctypes.data, xy.ctypes.data)
    got = [sum(int(v) << (64 * j) for j, v in enumerate(xy[6 * c:6 * c + 6])) for c in range(2)]
    assert got == [p[0], p[1]]
    q2 = kzg.g2_add(kzg.g2_mul(kzg.g2_generator(), k - 1), kzg.g2_generator())     # Z != 1
    xy2 = np.zeros(24, dtype=np.uint64)
    L.b200_g2_to_affine(np.ascontiguousarray(q2).ctypes.data, xy2.ctypes.data)
    got2 = [sum(int(v) << (64 * j) for j, v in enumerate(xy2[6 * c:6 * c + 6])) for c in range(4)]
    assert got2 == [q[0][0], q[0][1], q[1][0], q[1][1]]
    L.b200_g2_to_affine(np.zeros(36, dtype=np.uint64).ctypes.data, xy2.ctypes.data)
    assert not xy2.any()
