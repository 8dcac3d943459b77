"""Independent restatement of the BLS12-381 pairing with Python integers (TEST INFRASTRUCTURE, small cases only).

The product (go_kzg_b200/csrc/pairing.h) runs an affine Miller loop on the twist with sparse line values in
Fp2[w]/(w^6 - xi) and a split final exponentiation.  This file deliberately takes the other textbook route so that the
two share no formulas: Fp12 is Fp[w]/(w^12 - 2 w^6 + 2) (u = w^6 - 1), the G2 point is mapped to E(Fp12) first, the Miller
loop uses generic chord/tangent lines over Fp12 with polynomial inversion by the extended Euclidean algorithm, and the
final exponentiation is one plain power by (p^12 - 1) / r.  Both must produce the same GT element, coefficient for
coefficient.  Replaces nothing in the reference: kilic's pairing engine (bls/bls_kilic.go:152-158) is not in the tree.
"""
P = int("1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab", 16)
R = 52435875175126190479447740508185965837690552500527637822603658699938581184513
Z_ABS = 0xd201000000010000
G1 = (3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507,
      1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569)
G2 = ((352701069587466618187139116011060144890029952792775240219908644239793785735715026873347600343865175952761926303160,
       3059144344244213709971259814753781636986470325476647558659373206291635324768958432433509563104347017837885763365758),
      (1985150602287291935568054521177171638300868978215655730859378665066344726373823718423869104263333984641494340347905,
       927553665492332455747201965776037880757740193453592970025027978793976877002675564980949289727957565575433344219582))


# ------------------------------------------------------------------ Fp2 (pairs) and the curves over Fp / Fp2
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, -a[1] * n % P)


def g1_add(p, q):
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P == 0: return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, P) % P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x = (lam * lam - p[0] - q[0]) % P
    return (x, (lam * (p[0] - x) - p[1]) % P)


def g2_add(p, q):
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if f2_add(p[1], q[1]) == (0, 0): return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(p[0], p[0])), f2_inv(f2_add(p[1], p[1])))
    else:
        lam = f2_mul(f2_sub(q[1], p[1]), f2_inv(f2_sub(q[0], p[0])))
    x = f2_sub(f2_sub(f2_mul(lam, lam), p[0]), q[0])
    return (x, f2_sub(f2_mul(lam, f2_sub(p[0], x)), p[1]))


def _mul(add, p, k):
    acc = None
    while k:
        if k & 1: acc = add(acc, p)
        p = add(p, p); k >>= 1
    return acc


def g1_mul(p, k): return _mul(g1_add, p, k % R)
def g2_mul(p, k): return _mul(g2_add, p, k % R)


def g2_compress(p):
    """ZCash 96-byte form: x.c1 || x.c0 big-endian, flag bits 0x80 compressed, 0x40 infinity, 0x20 y is the larger root."""
    if p is None: return bytes([0xC0]) + bytes(95)
    (x0, x1), (y0, y1) = p
    big = y1 > (P - 1) // 2 if y1 else y0 > (P - 1) // 2
    b = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if big else 0)
    return bytes(b)


# ------------------------------------------------------------------ Fp12 = Fp[w] / (w^12 - 2 w^6 + 2)
MODPOLY = [2, 0, 0, 0, 0, 0, -2 % P, 0, 0, 0, 0, 0, 1]
ONE12 = [1] + [0] * 11


def f12_mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):          # w^12 = 2 w^6 - 2
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [v % P for v in t[:12]]


def f12_sub(a, b): return [(x - y) % P for x, y in zip(a, b)]
def f12_scalar(a, s): return [x * s % P for x in a]


def _trim(a):
    a = list(a)
    while a and a[-1] % P == 0: a.pop()
    return a


def _divmod(a, b):
    a = [v % P for v in a]
    q = [0] * max(1, len(a) - len(b) + 1)
    inv = pow(b[-1], -1, P)
    for k in range(len(a) - len(b), -1, -1):
        c = a[k + len(b) - 1] * inv % P
        q[k] = c
        if c:
            for j, bj in enumerate(b):
                a[k + j] = (a[k + j] - c * bj) % P
    return q, _trim(a[:len(b) - 1])


def _pmul(a, b):
    t = [0] * (len(a) + len(b) - 1) if a and b else []
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            t[i + j] = (t[i + j] + x * y) % P
    return t


def _psub(a, b):
    n = max(len(a), len(b))
    return _trim([((a[i] if i < len(a) else 0) - (b[i] if i < len(b) else 0)) % P for i in range(n)])


def f12_inv(a):
    r0, r1 = list(MODPOLY), _trim(a)
    s0, s1 = [], [1]
    assert r1, "zero has no inverse"
    while len(r1) > 1:
        q, r = _divmod(r0, r1)
        r0, r1 = r1, r
        s0, s1 = s1, _psub(s0, _pmul(q, s1))
    assert len(r1) == 1
    c = pow(r1[0], -1, P)
    out = [v * c % P for v in s1]
    return out + [0] * (12 - len(out))


def f12_pow(a, e):
    acc = list(ONE12)
    for bit in bin(e)[2:]:
        acc = f12_mul(acc, acc)
        if bit == "1": acc = f12_mul(acc, a)
    return acc


def f2_to_f12(a):                       # a0 + a1 u, u = w^6 - 1
    out = [0] * 12
    out[0] = (a[0] - a[1]) % P
    out[6] = a[1] % P
    return out


def fp_to_f12(a): return [a % P] + [0] * 11


W2_INV = None
W3_INV = None


def untwist(q):
    """E'(Fp2): y^2 = x^3 + 4(1 + u)  ->  E(Fp12): y^2 = x^3 + 4,  (x, y) -> (x / w^2, y / w^3)"""
    global W2_INV, W3_INV
    if W2_INV is None:
        W2_INV = f12_inv([0, 0, 1] + [0] * 9)
        W3_INV = f12_inv([0, 0, 0, 1] + [0] * 8)
    x, y = f12_mul(f2_to_f12(q[0]), W2_INV), f12_mul(f2_to_f12(q[1]), W3_INV)
    lhs = f12_mul(y, y)
    rhs = f12_mul(f12_mul(x, x), x)
    rhs[0] = (rhs[0] + 4) % P
    assert lhs == rhs
    return (x, y)


def _slope(p1, p2):
    if p1[0] != p2[0]:
        return f12_mul(f12_sub(p2[1], p1[1]), f12_inv(f12_sub(p2[0], p1[0])))
    assert p1[1] == p2[1]
    return f12_mul(f12_scalar(f12_mul(p1[0], p1[0]), 3), f12_inv(f12_scalar(p1[1], 2)))


def _line(p1, p2, t):
    m = _slope(p1, p2)
    return f12_sub(f12_mul(m, f12_sub(t[0], p1[0])), f12_sub(t[1], p1[1]))


def _ec12_add(p1, p2):
    m = _slope(p1, p2)
    x = f12_sub(f12_sub(f12_mul(m, m), p1[0]), p2[0])
    return (x, f12_sub(f12_mul(m, f12_sub(p1[0], x)), p1[1]))


def pairing(p, q):
    """e(P, Q) for affine P in G1 (pair of ints) and Q in G2 (pair of Fp2 pairs); None is infinity."""
    if p is None or q is None:
        return list(ONE12)
    q12 = untwist(q)
    p12 = (fp_to_f12(p[0]), fp_to_f12(p[1]))
    t, f = q12, list(ONE12)
    for bit in bin(Z_ABS)[3:]:
        f = f12_mul(f12_mul(f, f), _line(t, t, p12))
        t = _ec12_add(t, t)
        if bit == "1":
            f = f12_mul(f, _line(t, q12, p12))
            t = _ec12_add(t, q12)
    f = f12_inv(f)                      # z < 0
    return f12_pow(f, (P ** 12 - 1) // R)


def gt_from_flat(coeffs):
    """The product's layout (six Fp2 coefficients c_i = a_i + b_i u of w^i, as 12 integers a0, b0, a1, b1, ...) as a
    polynomial in w: c_i w^i = (a_i - b_i) w^i + b_i w^(i + 6)."""
    out = [0] * 12
    for i in range(6):
        a, b = coeffs[2 * i], coeffs[2 * i + 1]
        out[i] = (out[i] + a - b) % P
        out[i + 6] = (out[i + 6] + b) % P
    return out
