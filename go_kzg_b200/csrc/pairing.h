// pairing.h -- host-side G2 arithmetic and the optimal ate pairing of BLS12-381 (level-1 contract of the bls/ package:
// MulG2 / AddG2 / SubG2 / NegG2 / EqualG2 / To|FromCompressedG2 / PairingsVerify, bls/bls_kilic.go:69-104,123-130,152-158).
//
// Off the hot path: the verifier needs two G2 operations and one pairing check per proof (kzg_single_proofs.go:57-75,
// kzg_multi_proofs.go:47-88), so this is plain host C++ over the same Montgomery Fp as the device code, written for
// obviousness, not speed:
//     Fp2  = Fp[u] / (u^2 + 1)
//     Fp12 = Fp2[w] / (w^6 - xi), xi = 1 + u        (flat: six Fp2 coefficients; w^2 = v gives the usual Fp6 / Fp12 tower)
//     G2   = E'(Fp2)[r],  E': y^2 = x^3 + 4 xi       (M-type sextic twist; untwist (x, y) -> (x / w^2, y / w^3))
// Miller loop: f_{|z|,Q}(P) with |z| = 0xd201000000010000, conjugated because z < 0, T running in Jacobian coordinates on
// the twist; a line through T with slope lambda evaluated at P = (xP, yP) is, up to a factor in Fp4 that the final
// exponentiation removes,   (lambda x_T - y_T) - lambda xP w^2 + yP w^3   (slopes are never divided out: their
// denominators are in Fp2).
// Final exponentiation: (p^6 - 1)(p^2 + 1) by conjugation / Frobenius, then the hard part (p^4 - p^2 + 1) / r by
// square-and-multiply.  Checked against an independent Python restatement (tests/pairing_ref.py: Fp12 as polynomials
// modulo w^12 - 2 w^6 + 2, affine Miller loop on E(Fp12), plain power by (p^12 - 1) / r) and by bilinearity tests.
#pragma once
#include "hostutil.cuh"

namespace b200 {

// ---- Fp on the host with 64-bit limbs -------------------------------------------------------------------------------
// The HD routines of field.cuh run on the host as an emulation of the device's 32-bit carry chains (that is what the host
// tests of the device algorithms want, ~0.6 us per product); a pairing is ~3 * 10^5 products, so the code below multiplies
// with 6 x 64-bit limbs and unsigned __int128 instead.  Same Montgomery form (R = 2^384), same canonical results.
inline uint64_t fp_inv64() {             // -p^-1 mod 2^64 from the 32-bit constant by one Newton step
    const uint64_t p0 = (uint64_t)FpParams::mod(0) | ((uint64_t)FpParams::mod(1) << 32);
    uint64_t t = (uint64_t)(0u - FpParams::INV);      // p^-1 mod 2^32
    t *= 2 - p0 * t;                                    // mod 2^64
    return 0 - t;
}
struct FpMod64 { uint64_t m[6]; uint64_t inv; };
inline const FpMod64& fp_mod64c() {
    static const FpMod64 k = [] {
        FpMod64 r;
        for (int i = 0; i < 6; i++) r.m[i] = (uint64_t)FpParams::mod(2 * i) | ((uint64_t)FpParams::mod(2 * i + 1) << 32);
        r.inv = fp_inv64();
        return r;
    }();
    return k;
}
inline Fp hmul(const Fp& a, const Fp& b) {
    const FpMod64& K = fp_mod64c();
    const uint64_t* M = K.m;
    uint64_t A[6], B[6], t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    memcpy(A, a.l, 48); memcpy(B, b.l, 48);
    for (int i = 0; i < 6; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 6; j++) { c += (unsigned __int128)A[j] * B[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[6] = (uint64_t)c; t[7] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * K.inv;
        c = ((unsigned __int128)m * M[0] + t[0]) >> 64;
        for (int j = 1; j < 6; j++) { c += (unsigned __int128)m * M[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[6]; t[5] = (uint64_t)c; t[6] = t[7] + (uint64_t)(c >> 64);
    }
    // t < 2 p < 2^382 (t[6] == 0): one conditional subtraction
    uint64_t d[6], br = 0;
    for (int i = 0; i < 6; i++) { unsigned __int128 u = (unsigned __int128)t[i] - M[i] - br; d[i] = (uint64_t)u; br = (uint64_t)(u >> 64) & 1; }
    Fp r;
    memcpy(r.l, br ? t : d, 48);
    return r;
}
inline Fp hsqr(const Fp& a) { return hmul(a, a); }
template <int NE>
inline Fp hpow(const Fp& a, const uint32_t (&e)[NE]) {
    Fp acc = Fp::one();
    for (int i = NE * 32 - 1; i >= 0; i--) {
        acc = hsqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = hmul(acc, a);
    }
    return acc;
}
inline Fp hinv(const Fp& a) {            // Fermat; 0 -> 0
    uint32_t e[12];
    for (int i = 0; i < 12; i++) e[i] = FpParams::modm2(i);
    return hpow<12>(a, e);
}

inline Fp hadd(const Fp& a, const Fp& b) {      // (a + b) mod p, inputs < p
    uint64_t A[6], B[6], s[6], d[6];
    const uint64_t* M = fp_mod64c().m;
    memcpy(A, a.l, 48); memcpy(B, b.l, 48);
    unsigned __int128 c = 0;
    for (int i = 0; i < 6; i++) { c += (unsigned __int128)A[i] + B[i]; s[i] = (uint64_t)c; c >>= 64; }     // < 2^382: no carry out
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) { unsigned __int128 t = (unsigned __int128)s[i] - M[i] - br; d[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
    Fp r;
    memcpy(r.l, br ? s : d, 48);
    return r;
}
inline Fp hsub(const Fp& a, const Fp& b) {      // (a - b) mod p
    uint64_t A[6], B[6], d[6];
    const uint64_t* M = fp_mod64c().m;
    memcpy(A, a.l, 48); memcpy(B, b.l, 48);
    uint64_t br = 0;
    for (int i = 0; i < 6; i++) { unsigned __int128 t = (unsigned __int128)A[i] - B[i] - br; d[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 6; i++) { c += (unsigned __int128)d[i] + M[i]; d[i] = (uint64_t)c; c >>= 64; }
    }
    Fp r;
    memcpy(r.l, d, 48);
    return r;
}
inline Fp hneg(const Fp& a) { return hsub(Fp::zero(), a); }
inline Fp hdbl(const Fp& a) { return hadd(a, a); }

struct Fp2 { Fp c0, c1; };

inline Fp2 f2_zero() { Fp2 r; r.c0 = Fp::zero(); r.c1 = Fp::zero(); return r; }
inline Fp2 f2_one() { Fp2 r; r.c0 = Fp::one(); r.c1 = Fp::zero(); return r; }
inline bool f2_is_zero(const Fp2& a) { return a.c0.is_zero() && a.c1.is_zero(); }
inline bool f2_eq(const Fp2& a, const Fp2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
inline Fp2 f2_add(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = hadd(a.c0, b.c0); r.c1 = hadd(a.c1, b.c1); return r; }
inline Fp2 f2_sub(const Fp2& a, const Fp2& b) { Fp2 r; r.c0 = hsub(a.c0, b.c0); r.c1 = hsub(a.c1, b.c1); return r; }
inline Fp2 f2_neg(const Fp2& a) { Fp2 r; r.c0 = hneg(a.c0); r.c1 = hneg(a.c1); return r; }
inline Fp2 f2_dbl(const Fp2& a) { return f2_add(a, a); }
inline Fp2 f2_conj(const Fp2& a) { Fp2 r; r.c0 = a.c0; r.c1 = hneg(a.c1); return r; }
inline Fp2 f2_mul(const Fp2& a, const Fp2& b) {
    Fp t0 = hmul(a.c0, b.c0), t1 = hmul(a.c1, b.c1);
    Fp2 r;
    r.c0 = hsub(t0, t1);
    r.c1 = hsub(hsub(hmul(hadd(a.c0, a.c1), hadd(b.c0, b.c1)), t0), t1);
    return r;
}
inline Fp2 f2_sqr(const Fp2& a) {
    Fp2 r;
    r.c0 = hmul(hadd(a.c0, a.c1), hsub(a.c0, a.c1));
    r.c1 = hdbl(hmul(a.c0, a.c1));
    return r;
}
inline Fp2 f2_mul_fp(const Fp2& a, const Fp& s) { Fp2 r; r.c0 = hmul(a.c0, s); r.c1 = hmul(a.c1, s); return r; }
inline Fp2 f2_mul_xi(const Fp2& a) { Fp2 r; r.c0 = hsub(a.c0, a.c1); r.c1 = hadd(a.c0, a.c1); return r; }   // (1 + u) a
inline Fp2 f2_inv(const Fp2& a) {     // 0 -> 0
    Fp n = hadd(hsqr(a.c0), hsqr(a.c1));
    Fp ni = hinv(n);
    Fp2 r; r.c0 = hmul(a.c0, ni); r.c1 = hneg(hmul(a.c1, ni)); return r;
}
template <int NE>
inline Fp2 f2_pow(const Fp2& a, const uint32_t (&e)[NE]) {
    Fp2 acc = f2_one();
    for (int i = NE * 32 - 1; i >= 0; i--) {
        acc = f2_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = f2_mul(acc, a);
    }
    return acc;
}
// square root in Fp (p = 3 mod 4): v^((p+1)/4), false if v is not a square
inline bool fp_sqrt(Fp& out, const Fp& v) {
    constexpr uint32_t e[12] = B200_FP_SQRT_EXP;
    Fp y = hpow<12>(v, e);
    if (hsqr(y) != v) return false;
    out = y; return true;
}
// square root in Fp2 by the norm method; false if a is not a square
inline bool f2_sqrt(Fp2& out, const Fp2& a) {
    Fp2 r;
    if (a.c1.is_zero()) {
        Fp s;
        if (fp_sqrt(s, a.c0)) { r.c0 = s; r.c1 = Fp::zero(); }
        else if (fp_sqrt(s, hneg(a.c0))) { r.c0 = Fp::zero(); r.c1 = s; }     // (s u)^2 = -s^2
        else return false;
    } else {
        Fp n = hadd(hsqr(a.c0), hsqr(a.c1)), s;
        if (!fp_sqrt(s, n)) return false;
        Fp two = hdbl(Fp::one()), half = hinv(two);
        Fp t = hmul(hadd(a.c0, s), half), x0;
        if (!fp_sqrt(x0, t)) {
            t = hmul(hsub(a.c0, s), half);
            if (!fp_sqrt(x0, t)) return false;
        }
        if (x0.is_zero()) return false;
        r.c0 = x0;
        r.c1 = hmul(a.c1, hinv(hdbl(x0)));
    }
    if (!f2_eq(f2_sqr(r), a)) return false;
    out = r; return true;
}

// ABI point -> Montgomery Jacobian with the 64-bit product (g1_from_abi's fe_to_mont is the 32-bit emulation)
inline G1J g1_from_abi_h(const uint64_t* p) {
    G1J r;
    memcpy(r.x.l, p, 48); memcpy(r.y.l, p + 6, 48); memcpy(r.z.l, p + 12, 48);
    const Fp r2 = Fp::r2();
    r.x = hmul(r.x, r2); r.y = hmul(r.y, r2); r.z = hmul(r.z, r2);
    return r;
}
inline bool g1_on_curve(const G1J& p) {      // Y^2 = X^3 + 4 Z^6
    if (p.is_inf()) return true;
    Fp z2 = hsqr(p.z), z6 = hmul(hsqr(z2), z2);
    return hsqr(p.y) == hadd(hmul(hsqr(p.x), p.x), hmul(fp_const_four(), z6));
}

// ---------------------------------------------------------------------------------------------------------- G2
struct G2J {   // Jacobian over Fp2, Montgomery coordinates; infinity <=> z == 0
    Fp2 x, y, z;
    static G2J infinity() { G2J r; r.x = f2_zero(); r.y = f2_zero(); r.z = f2_zero(); return r; }
    bool is_inf() const { return f2_is_zero(z); }
};
inline Fp fp_from_limbs_canon(const uint32_t (&t)[12]) { Fp r; for (int i = 0; i < 12; i++) r.l[i] = t[i]; return fe_to_mont(r); }
inline Fp2 g2_curve_b() { Fp2 b; b.c0 = fp_const_four(); b.c1 = fp_const_four(); return b; }      // 4 (1 + u)
inline G2J g2_generator() {
    constexpr uint32_t x0[12] = B200_G2_GEN_X0, x1[12] = B200_G2_GEN_X1, y0[12] = B200_G2_GEN_Y0, y1[12] = B200_G2_GEN_Y1;
    G2J g;
    g.x.c0 = fp_from_limbs_canon(x0); g.x.c1 = fp_from_limbs_canon(x1);
    g.y.c0 = fp_from_limbs_canon(y0); g.y.c1 = fp_from_limbs_canon(y1);
    g.z = f2_one();
    return g;
}
inline G2J g2_neg(const G2J& p) { G2J r = p; r.y = f2_neg(p.y); return r; }
inline G2J g2_dbl(const G2J& p) {          // dbl-2009-l (a = 0)
    if (p.is_inf()) return p;
    Fp2 a = f2_sqr(p.x), b = f2_sqr(p.y), c = f2_sqr(b);
    Fp2 d = f2_dbl(f2_sub(f2_sub(f2_sqr(f2_add(p.x, b)), a), c));
    Fp2 e = f2_add(f2_dbl(a), a), f = f2_sqr(e);
    G2J r;
    r.z = f2_dbl(f2_mul(p.y, p.z));
    r.x = f2_sub(f, f2_dbl(d));
    r.y = f2_sub(f2_mul(e, f2_sub(d, r.x)), f2_dbl(f2_dbl(f2_dbl(c))));
    return r;
}
inline G2J g2_add(const G2J& p, const G2J& q) {   // add-2007-bl with the case analysis of the group law
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    Fp2 z1z1 = f2_sqr(p.z), z2z2 = f2_sqr(q.z);
    Fp2 u1 = f2_mul(p.x, z2z2), u2 = f2_mul(q.x, z1z1);
    Fp2 s1 = f2_mul(f2_mul(p.y, q.z), z2z2), s2 = f2_mul(f2_mul(q.y, p.z), z1z1);
    Fp2 h = f2_sub(u2, u1), rr = f2_sub(s2, s1);
    if (f2_is_zero(h)) return f2_is_zero(rr) ? g2_dbl(p) : G2J::infinity();
    Fp2 hh = f2_sqr(h), hhh = f2_mul(h, hh), v = f2_mul(u1, hh);
    G2J r;
    r.x = f2_sub(f2_sub(f2_sqr(rr), hhh), f2_dbl(v));
    r.y = f2_sub(f2_mul(rr, f2_sub(v, r.x)), f2_mul(s1, hhh));
    r.z = f2_mul(f2_mul(p.z, q.z), h);
    return r;
}
inline G2J g2_sub(const G2J& p, const G2J& q) { return g2_add(p, g2_neg(q)); }
// k * P, k as NL 32-bit limbs (canonical integer)
template <int NL>
inline G2J g2_mul_limbs(const G2J& p, const uint32_t (&k)[NL]) {
    G2J acc = G2J::infinity();
    for (int i = NL * 32 - 1; i >= 0; i--) {
        acc = g2_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) acc = g2_add(acc, p);
    }
    return acc;
}
inline G2J g2_mul(const G2J& p, const Fr& k_canon) {
    uint32_t k[8];
    for (int i = 0; i < 8; i++) k[i] = k_canon.l[i];
    return g2_mul_limbs<8>(p, k);
}
inline bool g2_equal(const G2J& p, const G2J& q) {
    if (p.is_inf() || q.is_inf()) return p.is_inf() && q.is_inf();
    Fp2 z1z1 = f2_sqr(p.z), z2z2 = f2_sqr(q.z);
    if (!f2_eq(f2_mul(p.x, z2z2), f2_mul(q.x, z1z1))) return false;
    return f2_eq(f2_mul(f2_mul(p.y, q.z), z2z2), f2_mul(f2_mul(q.y, p.z), z1z1));
}
inline bool g2_on_curve(const G2J& p) {      // Y^2 = X^3 + b Z^6
    if (p.is_inf()) return true;
    Fp2 z2 = f2_sqr(p.z), z6 = f2_mul(f2_sqr(z2), z2);
    return f2_eq(f2_sqr(p.y), f2_add(f2_mul(f2_sqr(p.x), p.x), f2_mul(g2_curve_b(), z6)));
}
inline bool g2_in_subgroup(const G2J& p) {   // r P == infinity
    uint32_t r[8];
    for (int i = 0; i < 8; i++) r[i] = FrParams::mod(i);
    return g2_mul_limbs<8>(p, r).is_inf();
}
inline void g2_affine(const G2J& p, Fp2& x, Fp2& y) {   // Montgomery affine coordinates of a finite point
    Fp2 zi = f2_inv(p.z), zi2 = f2_sqr(zi);
    x = f2_mul(p.x, zi2);
    y = f2_mul(p.y, f2_mul(zi2, zi));
}

// ABI: X, Y, Z each (c0, c1), every coefficient 6 x u64 canonical little-endian: 36 x u64 = 288 bytes; infinity <=> Z == 0
inline G2J g2_from_abi(const uint64_t* p) {
    G2J r;
    Fp* c[6] = {&r.x.c0, &r.x.c1, &r.y.c0, &r.y.c1, &r.z.c0, &r.z.c1};
    for (int i = 0; i < 6; i++) { memcpy(c[i]->l, p + 6 * i, 48); *c[i] = fe_to_mont(*c[i]); }
    return r;
}
inline void g2_to_abi(uint64_t* p, const G2J& a) {
    const Fp* c[6] = {&a.x.c0, &a.x.c1, &a.y.c0, &a.y.c1, &a.z.c0, &a.z.c1};
    for (int i = 0; i < 6; i++) { Fp v = fe_from_mont(*c[i]); memcpy(p + 6 * i, v.l, 48); }
}
inline bool fp_abi_canonical(const uint64_t* p) {      // value < p
    Fp x; memcpy(x.l, p, 48);
    for (int i = 11; i >= 0; i--) {
        if (x.l[i] < FpParams::mod(i)) return true;
        if (x.l[i] > FpParams::mod(i)) return false;
    }
    return false;
}

// ZCash 96-byte form: x.c1 || x.c0 big-endian, flags in the top three bits of byte 0 (compressed, infinity, y "largest":
// y.c1 > (p-1)/2, or y.c1 == 0 and y.c0 > (p-1)/2)
inline void fp_to_be48(uint8_t* out, const Fp& canon) {
    for (int i = 0; i < 48; i++) out[i] = (uint8_t)(canon.l[(47 - i) >> 2] >> (((47 - i) & 3) * 8));
}
inline bool fp_from_be48(Fp& canon, const uint8_t* in, uint8_t first_mask) {     // false if >= p
    canon = Fp::zero();
    for (int i = 0; i < 48; i++) {
        uint8_t b = in[i];
        if (i == 0) b &= first_mask;
        canon.l[(47 - i) >> 2] |= (uint32_t)b << (((47 - i) & 3) * 8);
    }
    for (int i = 11; i >= 0; i--) {
        if (canon.l[i] < FpParams::mod(i)) return true;
        if (canon.l[i] > FpParams::mod(i)) return false;
    }
    return false;
}
inline bool f2_y_is_largest(const Fp2& y_mont) {
    Fp c1 = fe_from_mont(y_mont.c1);
    if (!c1.is_zero()) return fp_canon_gt_half(c1);
    return fp_canon_gt_half(fe_from_mont(y_mont.c0));
}
inline void g2_compress(uint8_t out[96], const G2J& p) {
    memset(out, 0, 96);
    if (p.is_inf()) { out[0] = 0xC0; return; }
    Fp2 x, y;
    g2_affine(p, x, y);
    fp_to_be48(out, fe_from_mont(x.c1));
    fp_to_be48(out + 48, fe_from_mont(x.c0));
    out[0] |= 0x80;
    if (f2_y_is_largest(y)) out[0] |= 0x20;
}
// 0 ok, 1 malformed flags / coordinate >= p, 2 not on the curve, 3 not in the prime-order subgroup
inline int g2_decompress(G2J& p, const uint8_t in[96]) {
    if (!(in[0] & 0x80)) return 1;
    if (in[0] & 0x40) {
        for (int i = 1; i < 96; i++) if (in[i]) return 1;
        if (in[0] & 0x3F) return 1;
        p = G2J::infinity();
        return 0;
    }
    Fp c1, c0;
    if (!fp_from_be48(c1, in, 0x1F) || !fp_from_be48(c0, in + 48, 0xFF)) return 1;
    Fp2 x; x.c0 = fe_to_mont(c0); x.c1 = fe_to_mont(c1);
    Fp2 rhs = f2_add(f2_mul(f2_sqr(x), x), g2_curve_b()), y;
    if (!f2_sqrt(y, rhs)) return 2;
    if (f2_y_is_largest(y) != !!(in[0] & 0x20)) y = f2_neg(y);
    p.x = x; p.y = y; p.z = f2_one();
    if (!g2_in_subgroup(p)) return 3;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------- Fp12
struct Fp12 { Fp2 c[6]; };     // sum c[i] w^i, w^6 = xi

inline Fp12 f12_one() { Fp12 r; for (int i = 0; i < 6; i++) r.c[i] = f2_zero(); r.c[0] = f2_one(); return r; }
inline bool f12_eq(const Fp12& a, const Fp12& b) { for (int i = 0; i < 6; i++) if (!f2_eq(a.c[i], b.c[i])) return false; return true; }
// Fp6 = Fp2[v] / (v^3 - xi) (v = w^2): the even and the odd coefficients of an Fp12 element are Fp6 elements A0, A1 with
// a = A0 + A1 w, w^2 = v
struct Fp6 { Fp2 a, b, c; };
inline Fp6 f6_add(const Fp6& x, const Fp6& y) { Fp6 r; r.a = f2_add(x.a, y.a); r.b = f2_add(x.b, y.b); r.c = f2_add(x.c, y.c); return r; }
inline Fp6 f6_sub(const Fp6& x, const Fp6& y) { Fp6 r; r.a = f2_sub(x.a, y.a); r.b = f2_sub(x.b, y.b); r.c = f2_sub(x.c, y.c); return r; }
inline Fp6 f6_mul_v(const Fp6& x) { Fp6 r; r.a = f2_mul_xi(x.c); r.b = x.a; r.c = x.b; return r; }
inline Fp6 f6_mul(const Fp6& x, const Fp6& y) {       // Karatsuba: 6 Fp2 products
    Fp2 v0 = f2_mul(x.a, y.a), v1 = f2_mul(x.b, y.b), v2 = f2_mul(x.c, y.c);
    Fp6 r;
    r.a = f2_add(v0, f2_mul_xi(f2_sub(f2_sub(f2_mul(f2_add(x.b, x.c), f2_add(y.b, y.c)), v1), v2)));
    r.b = f2_add(f2_sub(f2_sub(f2_mul(f2_add(x.a, x.b), f2_add(y.a, y.b)), v0), v1), f2_mul_xi(v2));
    r.c = f2_add(f2_sub(f2_sub(f2_mul(f2_add(x.a, x.c), f2_add(y.a, y.c)), v0), v2), v1);
    return r;
}
inline void f12_split(const Fp12& a, Fp6& a0, Fp6& a1) { a0 = {a.c[0], a.c[2], a.c[4]}; a1 = {a.c[1], a.c[3], a.c[5]}; }
inline Fp12 f12_join(const Fp6& a0, const Fp6& a1) {
    Fp12 r;
    r.c[0] = a0.a; r.c[2] = a0.b; r.c[4] = a0.c;
    r.c[1] = a1.a; r.c[3] = a1.b; r.c[5] = a1.c;
    return r;
}
inline Fp12 f12_mul(const Fp12& a, const Fp12& b) {    // 3 Fp6 products
    Fp6 a0, a1, b0, b1;
    f12_split(a, a0, a1); f12_split(b, b0, b1);
    Fp6 t0 = f6_mul(a0, b0), t1 = f6_mul(a1, b1);
    Fp6 m = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1);
    return f12_join(f6_add(t0, f6_mul_v(t1)), m);
}
inline Fp12 f12_sqr(const Fp12& a) {                   // complex squaring: 2 Fp6 products
    Fp6 a0, a1;
    f12_split(a, a0, a1);
    Fp6 t = f6_mul(a0, a1);
    Fp6 c0 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(a0, f6_mul_v(a1))), t), f6_mul_v(t));
    return f12_join(c0, f6_add(t, t));
}
// a * b for a sparse a (the line values have three non-zero coefficients): schoolbook over the non-zero ones
inline Fp12 f12_mul_sparse(const Fp12& a, const Fp12& b) {
    Fp2 t[11];
    for (int i = 0; i < 11; i++) t[i] = f2_zero();
    for (int i = 0; i < 6; i++) {
        if (f2_is_zero(a.c[i])) continue;
        for (int j = 0; j < 6; j++) t[i + j] = f2_add(t[i + j], f2_mul(a.c[i], b.c[j]));
    }
    Fp12 r;
    for (int i = 0; i < 6; i++) r.c[i] = i < 5 ? f2_add(t[i], f2_mul_xi(t[i + 6])) : t[i];
    return r;
}
inline Fp12 f12_conj(const Fp12& a) {          // the p^6 Frobenius: w -> -w
    Fp12 r = a;
    r.c[1] = f2_neg(a.c[1]); r.c[3] = f2_neg(a.c[3]); r.c[5] = f2_neg(a.c[5]);
    return r;
}
// Frobenius x -> x^p: (c w^i)^p = conj(c) gamma^i w^i with gamma = w^(p-1) = xi^((p-1)/6)
inline Fp12 f12_frobenius(const Fp12& a) {
    static const Fp2 gamma = [] {
        constexpr uint32_t e[12] = B200_FP_PM1_DIV6;
        Fp2 xi; xi.c0 = Fp::one(); xi.c1 = Fp::one();
        return f2_pow<12>(xi, e);
    }();
    Fp12 r;
    Fp2 g = f2_one();
    for (int i = 0; i < 6; i++) { r.c[i] = f2_mul(f2_conj(a.c[i]), g); g = f2_mul(g, gamma); }
    return r;
}
inline Fp6 f6_inv(const Fp6& x) {
    Fp2 t0 = f2_sub(f2_sqr(x.a), f2_mul_xi(f2_mul(x.b, x.c)));
    Fp2 t1 = f2_sub(f2_mul_xi(f2_sqr(x.c)), f2_mul(x.a, x.b));
    Fp2 t2 = f2_sub(f2_sqr(x.b), f2_mul(x.a, x.c));
    Fp2 d = f2_add(f2_mul(x.a, t0), f2_mul_xi(f2_add(f2_mul(x.c, t1), f2_mul(x.b, t2))));
    Fp2 di = f2_inv(d);
    Fp6 r; r.a = f2_mul(t0, di); r.b = f2_mul(t1, di); r.c = f2_mul(t2, di);
    return r;
}
inline Fp12 f12_inv(const Fp12& a) {     // (A0 + A1 w)^-1 = (A0 - A1 w) / (A0^2 - v A1^2)
    Fp6 a0, a1;
    f12_split(a, a0, a1);
    Fp6 d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))));
    Fp6 r1 = f6_mul(a1, d);
    r1.a = f2_neg(r1.a); r1.b = f2_neg(r1.b); r1.c = f2_neg(r1.c);
    return f12_join(f6_mul(a0, d), r1);
}

// ---------------------------------------------------------------------------------------------------------- pairing
// Line values.  A line through T on the twist with slope lambda, evaluated at P = (xP, yP) in G1, is (see the header)
//     (lambda x_T - y_T) - lambda xP w^2 + yP w^3
// up to factors in Fp2, which the final exponentiation removes -- so the slopes never have to be divided out.  With T = (X, Y, Z)
// Jacobian on the twist (x = X / Z^2, y = Y / Z^3):
//   tangent at T:  lambda = 3 X^2 / (2 Y Z); times 2 Y Z^3:   (3 X^3 - 2 Y^2)  -  3 X^2 Z^2 xP w^2  +  2 Y Z^3 yP w^3
//   chord T, Q (Q affine): lambda = r / (Z h), r = y_Q Z^3 - Y, h = x_Q Z^2 - X; times Z h, through Q:
//                                                               (r x_Q - y_Q Z h)  -  r xP w^2  +  Z h yP w^3
struct MillerPair {
    Fp px, py;          // P affine
    Fp2 qx, qy;         // Q affine
    G2J t;              // running multiple of Q
};
inline Fp12 miller_sparse(const Fp2& c0, const Fp2& c2, const Fp2& c3) {
    Fp12 l;
    for (int i = 0; i < 6; i++) l.c[i] = f2_zero();
    l.c[0] = c0; l.c[2] = c2; l.c[3] = c3;
    return l;
}
inline Fp12 miller_double_step(MillerPair& m) {
    const G2J& t = m.t;
    Fp2 x2 = f2_sqr(t.x), y2 = f2_sqr(t.y), z2 = f2_sqr(t.z);
    Fp2 x2_3 = f2_add(f2_dbl(x2), x2);
    Fp2 c0 = f2_sub(f2_mul(x2_3, t.x), f2_dbl(y2));
    Fp2 c2 = f2_neg(f2_mul_fp(f2_mul(x2_3, z2), m.px));
    Fp2 c3 = f2_mul_fp(f2_mul(f2_dbl(f2_mul(t.y, t.z)), z2), m.py);
    m.t = g2_dbl(t);
    return miller_sparse(c0, c2, c3);
}
inline Fp12 miller_add_step(MillerPair& m) {
    const G2J& t = m.t;
    Fp2 z2 = f2_sqr(t.z);
    Fp2 r = f2_sub(f2_mul(m.qy, f2_mul(z2, t.z)), t.y), h = f2_sub(f2_mul(m.qx, z2), t.x);
    Fp2 zh = f2_mul(t.z, h);
    Fp2 c0 = f2_sub(f2_mul(r, m.qx), f2_mul(m.qy, zh));
    Fp2 c2 = f2_neg(f2_mul_fp(r, m.px));
    Fp2 c3 = f2_mul_fp(zh, m.py);
    G2J q; q.x = m.qx; q.y = m.qy; q.z = f2_one();
    m.t = g2_add(t, q);               // T = kQ with 1 < k < r - 1: never +-Q, so h != 0
    return miller_sparse(c0, c2, c3);
}
// prod_i f_{|z|,Q_i}(P_i), conjugated because z < 0; pairs with an infinite point contribute 1.  The squarings of the
// accumulator are shared between the pairs.
inline Fp12 miller_loop_multi(const G1J* ps, const G2J* qs, int n) {
    MillerPair pairs[4];
    int m = 0;
    for (int i = 0; i < n && m < 4; i++) {
        if (ps[i].is_inf() || qs[i].is_inf()) continue;
        Fp zi = hinv(ps[i].z), zi2 = hsqr(zi);
        pairs[m].px = hmul(ps[i].x, zi2); pairs[m].py = hmul(ps[i].y, hmul(zi2, zi));
        g2_affine(qs[i], pairs[m].qx, pairs[m].qy);
        pairs[m].t.x = pairs[m].qx; pairs[m].t.y = pairs[m].qy; pairs[m].t.z = f2_one();
        m++;
    }
    Fp12 f = f12_one();
    if (m == 0) return f;
    const uint64_t z = B200_BLS_Z_ABS;
    for (int bit = 62; bit >= 0; bit--) {
        f = f12_sqr(f);
        for (int i = 0; i < m; i++) f = f12_mul_sparse(miller_double_step(pairs[i]), f);
        if ((z >> bit) & 1)
            for (int i = 0; i < m; i++) f = f12_mul_sparse(miller_add_step(pairs[i]), f);
    }
    return f12_conj(f);
}
inline Fp12 miller_loop(const G1J& p, const G2J& q) { return miller_loop_multi(&p, &q, 1); }
inline Fp12 final_exponentiation(const Fp12& f) {
    Fp12 t = f12_mul(f12_conj(f), f12_inv(f));                    // f^(p^6 - 1)
    t = f12_mul(f12_frobenius(f12_frobenius(t)), t);              // ^(p^2 + 1)
    constexpr uint32_t e[B200_PAIRING_HARD_LIMBS] = B200_PAIRING_HARD_EXP;   // (p^4 - p^2 + 1) / r
    Fp12 acc = f12_one();
    bool started = false;
    for (int i = B200_PAIRING_HARD_LIMBS * 32 - 1; i >= 0; i--) {
        if (started) acc = f12_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) { acc = started ? f12_mul(acc, t) : t; started = true; }
    }
    return acc;
}
inline Fp12 pairing(const G1J& p, const G2J& q) { return final_exponentiation(miller_loop(p, q)); }

// f^|z| for f in the cyclotomic subgroup (after the easy part)
inline Fp12 f12_pow_z_abs(const Fp12& f) {
    const uint64_t z = B200_BLS_Z_ABS;
    Fp12 acc = f;
    for (int bit = 62; bit >= 0; bit--) {
        acc = f12_sqr(acc);
        if ((z >> bit) & 1) acc = f12_mul(acc, f);
    }
    return acc;
}
// Is final_exponentiation(f) == 1?  Uses 3 (p^4 - p^2 + 1) / r = (z - 1)^2 (z + p) (z^2 + p^2 - 1) + 3 (checked in
// tools/gen_constants.py): five powers by |z| instead of a 1269-bit exponent.  It computes the CUBE of the pairing value,
// which is 1 exactly when the value is (3 does not divide r), so this form serves the check only.  After the easy part the
// element is unitary: inverse = conjugate, and g^z = conj(g^|z|) because z < 0.
inline bool final_exponentiation_is_one(const Fp12& f) {
    Fp12 t = f12_mul(f12_conj(f), f12_inv(f));
    t = f12_mul(f12_frobenius(f12_frobenius(t)), t);
    auto pow_z = [](const Fp12& g) { return f12_conj(f12_pow_z_abs(g)); };
    Fp12 a = f12_mul(pow_z(t), f12_conj(t));                                     // t^(z - 1)
    a = f12_mul(pow_z(a), f12_conj(a));                                          // t^((z - 1)^2)
    Fp12 b = f12_mul(pow_z(a), f12_frobenius(a));                                // ^(z + p)
    Fp12 c = f12_mul(f12_mul(pow_z(pow_z(b)), f12_frobenius(f12_frobenius(b))), f12_conj(b));   // ^(z^2 + p^2 - 1)
    c = f12_mul(c, f12_mul(f12_sqr(t), t));                                      // * t^3
    return f12_eq(c, f12_one());
}
// e(a1, a2) == e(b1, b2)   (bls/bls_kilic.go:152-158: e(a1^-1, a2) e(b1, b2) == 1)
inline bool pairings_verify(const G1J& a1, const G2J& a2, const G1J& b1, const G2J& b2) {
    const G1J ps[2] = {g1_neg(a1), b1};
    const G2J qs[2] = {a2, b2};
    return final_exponentiation_is_one(miller_loop_multi(ps, qs, 2));
}

}  // namespace b200
