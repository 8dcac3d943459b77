"""Pure-Python-int restatement of go-kzg's hot path (TEST INFRASTRUCTURE ONLY).

This module is the slow, obviously-correct half of the oracle.  It restates the
reference's algorithms with Python integers (semantics of `bls/bignum_pure.go`:
canonical residues mod r) and a textbook BLS12-381 G1 (affine, Python ints).
It exists to (1) pin the C oracle (`oracle/kzg_oracle.c`) and (2) be checked
itself against every golden vector the reference's tests hold
(`tests/golden/reference_goldens.json`, produced by `tests/golden/make_golden.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may
import anything under `oracle/`.  Nothing in `go_kzg_b200/` imports it.

The arithmetic itself lives in the un-vendored dependency
`github.com/kilic/bls12-381 v0.1.1-0.20220929213557-ca162e8a70f4` (go.mod:8);
we restate its *published* behaviour (field mod r / mod p, short-Weierstrass
y^2 = x^3 + 4, ZCash-style 48-byte compression) and anchor on the reference's
call sites and golden vectors.  All `file:line` citations are relative to
/root/reference.
"""
from __future__ import annotations

# ----------------------------------------------------------------------------
# Fields (bls/globals.go:9 ModulusStr; base prime is the BLS12-381 standard)
# ----------------------------------------------------------------------------
R_MOD = 52435875175126190479447740508185965837690552500527637822603658699938581184513
P_MOD = int(
    "1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f624"
    "1eabfffeb153ffffb9feffffffffaaab", 16)
PRIMITIVE_ROOT = 7  # bls/globals.go:27-60 roots are 7^((r-1)/2^k)

# generator of G1 (decimal coordinates as in bls/bls_hbls.go:23-24)
G1_X = 3685416753713387016781088315183077757961620795782546409894578378688607592378376318836054947676345821548104185464507
G1_Y = 1339506544944476473020471379941921221584933875938349620426543736416511423956333506472724655353366534992391756441569


def scale2_root_of_unity(k: int) -> int:
    """bls/globals.go:27-60: entry k of Scale2RootOfUnity = 7^((r-1)/2^k)."""
    return pow(PRIMITIVE_ROOT, (R_MOD - 1) >> k, R_MOD)


def inv_fr(a: int) -> int:
    return pow(a, R_MOD - 2, R_MOD)


# ----------------------------------------------------------------------------
# reverse_bit_order.go:81-101
# ----------------------------------------------------------------------------
def reverse_bits_limited(length: int, value: int) -> int:
    """reverse_bit_order.go:81-84"""
    bits = length.bit_length() - 1
    out = 0
    for i in range(bits):
        if value >> i & 1:
            out |= 1 << (bits - 1 - i)
    return out


def reverse_bit_order(values: list) -> None:
    """reverse_bit_order.go:86-101 (in place)."""
    n = len(values)
    assert n & (n - 1) == 0
    for i in range(n):
        r = reverse_bits_limited(n, i)
        if r > i:
            values[i], values[r] = values[r], values[i]


# ----------------------------------------------------------------------------
# fft.go:21-61
# ----------------------------------------------------------------------------
class FFTSettings:
    def __init__(self, max_scale: int):
        """fft.go:44-61 NewFFTSettings; fft.go:21-32 expandRootOfUnity."""
        self.max_width = 1 << max_scale
        self.root_of_unity = scale2_root_of_unity(max_scale)
        rootz = [1, self.root_of_unity]
        while rootz[-1] != 1:
            rootz.append(rootz[-1] * self.root_of_unity % R_MOD)
        self.expanded = rootz                 # W+1 entries, first == last == 1
        self.reverse = rootz[::-1]

    # ---- fft_fr.go:8-53 ----------------------------------------------------
    def _simple_ft(self, vals, off, stride, roots, rstride, l):
        out = []
        for i in range(l):
            last = vals[off] * roots[0] % R_MOD
            for j in range(1, l):
                v = vals[off + j * stride] * roots[((i * j) % l) * rstride] % R_MOD
                last = (last + v) % R_MOD
            out.append(last)
        return out

    def _fft(self, vals, off, stride, roots, rstride, l):
        if l <= 4:
            return self._simple_ft(vals, off, stride, roots, rstride, l)
        half = l >> 1
        L = self._fft(vals, off, stride << 1, roots, rstride << 1, half)
        R = self._fft(vals, off + stride, stride << 1, roots, rstride << 1, half)
        out = [0] * l
        for i in range(half):
            t = R[i] * roots[i * rstride] % R_MOD
            out[i] = (L[i] + t) % R_MOD
            out[i + half] = (L[i] - t) % R_MOD
        return out

    def fft(self, vals, inv=False):
        """fft_fr.go:55-105 FFT + InplaceFFT (zero-pad to pow2, natural order)."""
        n = len(vals)
        if n > self.max_width:
            raise ValueError("got %d values but only have %d roots of unity" % (n, self.max_width))
        n2 = 1 if n == 0 else 1 << (n - 1).bit_length()
        vals = list(vals) + [0] * (n2 - n)
        n = n2
        stride = self.max_width // n
        if inv:
            out = self._fft(vals, 0, 1, self.reverse, stride, n)
            ninv = inv_fr(n)
            return [x * ninv % R_MOD for x in out]
        return self._fft(vals, 0, 1, self.expanded, stride, n)

    # ---- das_extension.go:7-84 --------------------------------------------
    def _das_ext(self, ab, stride):
        n = len(ab)
        if n == 2:
            x = (ab[0] + ab[1]) % R_MOD
            y = (ab[0] - ab[1]) % R_MOD
            t = y * self.expanded[stride] % R_MOD
            return [(x + t) % R_MOD, (x - t) % R_MOD]
        hh = n >> 1
        a0 = [0] * hh
        a1 = [0] * hh
        for i in range(hh):
            a0[i] = (ab[i] + ab[hh + i]) % R_MOD
            a1[i] = (ab[i] - ab[hh + i]) * self.reverse[i * 2 * stride] % R_MOD
        a0 = self._das_ext(a0, stride << 1)
        a1 = self._das_ext(a1, stride << 1)
        o0 = [0] * hh
        o1 = [0] * hh
        for i in range(hh):
            t = a1[i] * self.expanded[(1 + 2 * i) * stride] % R_MOD
            o0[i] = (a0[i] + t) % R_MOD
            o1[i] = (a0[i] - t) % R_MOD
        return o0 + o1

    def das_fft_extension(self, vals):
        """das_extension.go:71-84 (returns the odd samples; the Go mutates in place)."""
        if len(vals) * 2 > self.max_width:
            raise RuntimeError("domain too small for extending requested values")
        out = self._das_ext(list(vals), 1)
        ninv = inv_fr(len(vals))
        return [x * ninv % R_MOD for x in out]

    # ---- zero_poly.go ------------------------------------------------------
    def make_zero_poly_mul_leaf(self, dst_len, indices, stride):
        """zero_poly.go:17-39"""
        assert dst_len >= len(indices) + 1
        dst = [0] * dst_len
        dst[len(indices)] = 1
        for i, v in enumerate(indices):
            neg = (-self.expanded[v * stride]) % R_MOD
            dst[i] = neg
            if i > 0:
                dst[i] = (dst[i] + dst[i - 1]) % R_MOD
                for j in range(i - 1, 0, -1):
                    dst[j] = (dst[j] * neg + dst[j - 1]) % R_MOD
                dst[0] = dst[0] * neg % R_MOD
        return dst

    def reduce_leaves(self, n, ps):
        """zero_poly.go:58-107: product of <=4 polys via FFT of size n."""
        out_degree = sum(len(p) - 1 for p in ps)
        assert out_degree + 1 <= n
        # NOTE zero_poly.go:87,94-96: pPadded is only partially overwritten for
        # the non-last leaves (stale tail of the *last* leaf can remain when an
        # earlier leaf is shorter than the last one).  Restated faithfully.
        p_padded = list(ps[-1]) + [0] * (n - len(ps[-1]))
        mul_eval = self.fft_exact(p_padded, False)
        for p in ps[:-1]:
            for j, c in enumerate(p):
                p_padded[j] = c
            ev = self.fft_exact(p_padded, False)
            mul_eval = [a * b % R_MOD for a, b in zip(mul_eval, ev)]
        return self.fft_exact(mul_eval, True)[: out_degree + 1]

    def fft_exact(self, vals, inv):
        n = len(vals)
        assert n & (n - 1) == 0 and n <= self.max_width
        return self.fft(vals, inv)

    def zero_poly_via_multiplication(self, missing, length):
        """zero_poly.go:116-217 -> (zeroEval, zeroPoly)."""
        if len(missing) == 0:
            return [0] * length, [0] * length
        if length > self.max_width:
            raise RuntimeError("domain too small for requested length")
        assert length & (length - 1) == 0
        stride = self.max_width // length
        per_leaf_poly = 64
        per_leaf = per_leaf_poly - 1
        if len(missing) <= per_leaf:
            zp = self.make_zero_poly_mul_leaf(len(missing) + 1, missing, stride)
            zp = zp + [0] * (length - len(zp))
            return self.fft(zp, False), zp
        leaf_count = (len(missing) + per_leaf - 1) // per_leaf
        n = 1 << (leaf_count * per_leaf_poly - 1).bit_length()
        out = [0] * n
        # leaves are (offset, length) views into `out`, as in the Go slices
        leaves = []
        off = 0
        for i in range(leaf_count):
            idx = missing[off: off + per_leaf]
            leaf = self.make_zero_poly_mul_leaf(per_leaf_poly, idx, stride)
            out[i * per_leaf_poly:(i + 1) * per_leaf_poly] = leaf
            leaves.append((i * per_leaf_poly, per_leaf_poly))
            off += per_leaf
        rf = 4
        while len(leaves) > 1:
            reduced_count = (len(leaves) + rf - 1) // rf
            leaf_size = 1 << (leaves[0][1] - 1).bit_length()
            for i in range(reduced_count):
                start = i * rf
                end = start + rf
                out_end = min(end * leaf_size, len(out))
                red_off, red_len = start * leaf_size, out_end - start * leaf_size
                end = min(end, len(leaves))
                sl = leaves[start:end]
                if end > start + 1:
                    ps = [out[o:o + l] for (o, l) in sl]
                    res = self.reduce_leaves(red_len, ps)
                    # InplaceFFT writes all red_len entries of dst (zero_poly.go:103)
                    full = res + [0] * (red_len - len(res))
                    # (the inverse FFT output beyond outDegree is exactly zero)
                    out[red_off:red_off + red_len] = full
                    leaves[i] = (red_off, len(res))
                else:
                    # zero_poly.go:192,199: a lone leaf is *re-sliced* to the
                    # whole reduced window, not copied.
                    leaves[i] = (red_off, red_len)
            leaves = leaves[:reduced_count]
        o, l = leaves[0]
        zp = out[o:o + l]
        if l < length:
            zp = zp + [0] * (length - l)
        elif l > length:
            raise RuntimeError("expected output smaller or equal to input length")
        return self.fft(zp, False), zp

    # ---- recover_from_samples.go ------------------------------------------
    @staticmethod
    def shift_poly(poly):
        """recover_from_samples.go:9-24 (multiply coeff i by 5^-i)."""
        f = inv_fr(5)
        acc = 1
        out = []
        for c in poly:
            out.append(c * acc % R_MOD)
            acc = acc * f % R_MOD
        return out

    @staticmethod
    def unshift_poly(poly):
        """recover_from_samples.go:27-40"""
        acc = 1
        out = []
        for c in poly:
            out.append(c * acc % R_MOD)
            acc = acc * 5 % R_MOD
        return out

    def recover_poly_from_samples(self, samples):
        """recover_from_samples.go:42-109; samples: list of int or None."""
        n = len(samples)
        missing = [i for i, s in enumerate(samples) if s is None]
        zero_eval, zero_poly = self.zero_poly_via_multiplication(missing, n)
        for i, s in enumerate(samples):
            if (s is None) != (zero_eval[i] == 0):
                raise RuntimeError("bad zero eval")
        e = [0 if s is None else s * zero_eval[i] % R_MOD for i, s in enumerate(samples)]
        pwz = self.fft(e, True)
        spwz = self.shift_poly(pwz)
        szp = self.shift_poly(zero_poly)
        a = self.fft(spwz, False)
        b = self.fft(szp, False)
        q = [x * inv_fr(y) % R_MOD for x, y in zip(a, b)]
        srp = self.fft(q, True)
        rp = self.unshift_poly(srp)
        data = self.fft(rp, False)
        for i, s in enumerate(samples):
            if s is not None and data[i] != s:
                raise ValueError("failed to reconstruct data correctly, changed value at index %d" % i)
        return data


# ----------------------------------------------------------------------------
# G1: textbook affine arithmetic on y^2 = x^3 + 4 over Fp (None = infinity)
# ----------------------------------------------------------------------------
def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % P_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, P_MOD - 2, P_MOD) % P_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, P_MOD - 2, P_MOD) % P_MOD
    x3 = (lam * lam - x1 - x2) % P_MOD
    return x3, (lam * (x1 - x3) - y1) % P_MOD


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % P_MOD)


def g1_mul(P, k: int):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


G1_GEN = (G1_X, G1_Y)


def g1_compress(P) -> bytes:
    """48-byte ZCash/IETF form (pinned by bls/bls_test.go:18 and the JSON setup)."""
    if P is None:
        return bytes([0xC0]) + bytes(47)
    x, y = P
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80
    if y > (P_MOD - 1) // 2:
        b[0] |= 0x20
    return bytes(b)


def g1_decompress(b: bytes):
    assert len(b) == 48 and b[0] & 0x80
    if b[0] & 0x40:
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    y = pow((x * x * x + 4) % P_MOD, (P_MOD + 1) // 4, P_MOD)
    assert (y * y - x * x * x - 4) % P_MOD == 0, "not on curve"
    if (y > (P_MOD - 1) // 2) != bool(b[0] & 0x20):
        y = P_MOD - y
    return x, y


# ----------------------------------------------------------------------------
# Exponent-domain KZG oracle (setup secret known => every G1 output is k*G)
# ----------------------------------------------------------------------------
def eval_poly(coeffs, x):
    """bls/globals.go:76-95 Horner."""
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R_MOD
    return acc


def fk20_single_exponents(poly, secret, n2_domain_scale=None):
    """Discrete logs of FK20Single(poly) (fk20_single.go:122-134): proof i is
    q_i(s)*G with q_i = (p(X) - p(w^i)) / (X - w^i), w = root of unity of order
    len(poly), natural order."""
    n = len(poly)
    scale = n.bit_length() - 1
    w = scale2_root_of_unity(scale)
    ps = eval_poly(poly, secret)
    out = []
    x = 1
    for _ in range(n):
        y = eval_poly(poly, x)
        out.append((ps - y) * inv_fr((secret - x) % R_MOD) % R_MOD)
        x = x * w % R_MOD
    return out


def toeplitz_coeffs_step(poly):
    """fk20_single.go:106-119"""
    n = len(poly)
    return [poly[n - 1]] + [0] * (n + 1) + list(poly[1:n - 1])


def toeplitz_coeffs_step_strided(poly, offset, stride):
    """fk20_single.go:89-103"""
    n = len(poly)
    k = n // stride
    out = [0] * (2 * k)
    out[0] = poly[n - 1 - offset]
    j = 2 * stride - offset - 1
    for i in range(k + 2, 2 * k):
        out[i] = poly[j]
        j += stride
    return out


def fk20_pipeline_exponents(fs: FFTSettings, poly, secret, da: bool):
    """Restates FK20Single / DAUsingFK20 (fk20_single.go:122-196) entirely in
    the exponent: a G1 point k*G is represented by k.  Returns discrete logs."""
    n = len(poly)
    x = [pow(secret, j, R_MOD) for j in range(n - 2, -1, -1)] + [0]   # kzg.go:57-61
    x_ext_fft = fs.fft(x + [0] * n, False)                            # fk20_single.go:40-56
    c = fs.fft(toeplitz_coeffs_step(poly), False)                      # :59-66
    h_ext_fft = [a * b % R_MOD for a, b in zip(c, x_ext_fft)]         # :72-74
    h = fs.fft(h_ext_fft, True)[:n]                                   # :80-87
    if da:
        out = fs.fft(h + [0] * n, False)                              # :163-167
        reverse_bit_order(out)                                        # :194
        return out
    return fs.fft(h, False)


def fk20_multi_da_exponents(fs: FFTSettings, poly, secret, chunk_len):
    """Restates DAUsingFK20Multi (fk20_multi.go:58-133, kzg.go:73-116) in the
    exponent."""
    n = len(poly)
    k = n // chunk_len
    h_ext_fft = [0] * (2 * k)
    for off in range(chunk_len):
        start = n - chunk_len - 1 - off
        x = [pow(secret, start - i * chunk_len, R_MOD) for i in range(k - 1)] + [0]
        x_ext_fft = fs.fft(x + [0] * k, False)
        c = fs.fft(toeplitz_coeffs_step_strided(poly, off, chunk_len), False)
        for j in range(2 * k):
            h_ext_fft[j] = (h_ext_fft[j] + c[j] * x_ext_fft[j]) % R_MOD
    h = fs.fft(h_ext_fft, True)[:k]
    out = fs.fft(h + [0] * k, False)
    reverse_bit_order(out)
    return out


# ----------------------------------------------------------------------------
# package eth: the callers either side of the path
# ----------------------------------------------------------------------------
def valid_fr(le32: bytes) -> bool:
    """bls/bignum_all.go:12-35 ValidFr: a little-endian uint256 is a field element iff it is < r."""
    return int.from_bytes(le32, "little") < R_MOD


def blob_to_polynomial(blob: bytes):
    """eth/helpers.go:264-273 BlobToPolynomial: 32 little-endian bytes per element (bytesToBLSField ->
    bls.FrFrom32, eth/helpers.go:105-109); returns (elements, ok), ok False if any element is >= r."""
    out = []
    for i in range(0, len(blob), 32):
        chunk = blob[i:i + 32]
        if not valid_fr(chunk):
            return [], False
        out.append(int.from_bytes(chunk, "little"))
    return out, True


def eth_domain(width: int):
    """eth/globals.go:53-67: DomainFr[i] = ROOT_OF_UNITY^reverseBits(i), ROOT_OF_UNITY = 7^((r-1)/width)."""
    w = pow(PRIMITIVE_ROOT, (R_MOD - 1) // width, R_MOD)
    return [pow(w, reverse_bits_limited(width, i), R_MOD) for i in range(width)]


def bit_reversal_permutation(values):
    """eth/helpers.go:41-51: out[i] = l[reverseBits(i)] (applied to setup_G1_lagrange at eth/globals.go:48)."""
    n = len(values)
    return [values[reverse_bits_limited(n, i)] for i in range(n)]


def evaluate_poly_in_evaluation_form(poly, x, roots, scale: int = 0):
    """bls/globals.go:106-153 EvaluatePolyInEvaluationForm (barycentric formula; x outside the domain):
    y = (x^n - 1) / n * sum_i poly[i] roots[i << scale] / (x - roots[i << scale])."""
    n = len(poly)
    if n != len(roots) >> scale:
        raise ValueError("expected roots of unity to match polynomial size")      # bls/globals.go:107-109
    inv_denom = [(x - roots[i << scale]) % R_MOD for i in range(n)]
    inv_denom = [inv_fr(d) for d in inv_denom]                                   # BatchInvModFr, bls/globals.go:118-124
    y = 0
    for i in range(n):
        y = (y + poly[i] * roots[i << scale] % R_MOD * inv_denom[i]) % R_MOD     # bls/globals.go:127-141
    pow_b = (pow(x, n, R_MOD) - 1) % R_MOD                                       # bls/globals.go:143-146
    return y * pow_b % R_MOD * inv_fr(n) % R_MOD                                 # bls/globals.go:147-152


def compute_kzg_proof_quotient(poly, z, domain):
    """eth/helpers.go:179-203 ComputeKZGProof, field side: returns (y, quotient in evaluation form); the proof is
    LinCombG1(kzgSetupLagrange, quotient).  Raises for "invalid z challenge" (z in the domain) and for a
    polynomial whose length differs from the domain's."""
    if len(poly) != len(domain):
        raise ValueError("polynomial has invalid length")                        # eth/helpers.go:186-188
    y = evaluate_poly_in_evaluation_form(poly, z, domain, 0)                      # eth/helpers.go:180
    q = []
    for i in range(len(poly)):
        if domain[i] == z:
            raise ValueError("invalid z challenge")                              # eth/helpers.go:190-192
        num = (poly[i] - y) % R_MOD                                              # eth/helpers.go:182-184
        den = (domain[i] - z) % R_MOD                                            # eth/helpers.go:193
        q.append(num * inv_fr(den) % R_MOD)                                      # eth/helpers.go:196-198 DivModFr
    return y, q
