// g1.cuh -- BLS12-381 G1 in Jacobian coordinates over the Montgomery Fp of field.cuh.
//
// Group law is complete by case analysis (infinity, P == Q, P == -Q): FK20 zero-pads half of
// every G1 transform with infinity (fk20_single.go:48-50,163-166; kzg.go:61) and structured
// inputs do hit the doubling / cancellation branches.
//
// Replaces (semantics): kilic PointG1 Add/Sub/Double/MulScalar reached through
// bls/bls_kilic.go:41-65 (MulG1/AddG1/SubG1/NegG1).
#pragma once
#include "field.cuh"

namespace b200 {

struct G1J {   // Jacobian, Montgomery coordinates; infinity <=> z == 0
    Fp x, y, z;
    static HD G1J infinity() { G1J r; r.x = Fp::zero(); r.y = Fp::zero(); r.z = Fp::zero(); return r; }
    HD bool is_inf() const { return z.is_zero(); }
};
struct G1A {   // affine, Montgomery coordinates; infinity <=> (0,0) (not on y^2 = x^3 + 4)
    Fp x, y;
    HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

HD Fp fp_const_beta() { Fp r; constexpr uint32_t t[12] = B200_GLV_BETA; for (int i = 0; i < 12; i++) r.l[i] = t[i]; return r; }
HD Fp fp_const_four() { Fp r; constexpr uint32_t t[12] = B200_FP_FOUR; for (int i = 0; i < 12; i++) r.l[i] = t[i]; return r; }
HD G1J g1_generator() {
    G1J g;
    constexpr uint32_t gx[12] = B200_G1_GEN_X; constexpr uint32_t gy[12] = B200_G1_GEN_Y;
    for (int i = 0; i < 12; i++) { g.x.l[i] = gx[i]; g.y.l[i] = gy[i]; }
    g.z = Fp::one();
    return g;
}

HD G1J g1_neg(const G1J& p) { G1J r = p; r.y = fe_neg(p.y); return r; }

// 2P: 2M + 5S (a = 0; dbl-2009-l: 4 X Y^2 is taken from (X + Y^2)^2, squares being cheaper than products)
HD G1J g1_dbl(const G1J& p) {
    if (p.is_inf()) return p;
    Fp a = fp_sqr(p.x);
    Fp b = fp_sqr(p.y);
    Fp c = fp_sqr(b);
    Fp d = fe_sub(fe_sub(fp_sqr(fe_add(p.x, b)), a), c);
    d = fe_dbl(d);                    // 4 X Y^2
    Fp e = fe_add(fe_dbl(a), a);      // 3 X^2
    Fp f = fp_sqr(e);
    G1J r;
    r.z = fe_dbl(fp_mul(p.y, p.z));
    r.x = fe_sub(f, fe_dbl(d));
    Fp c8 = fe_dbl(fe_dbl(fe_dbl(c)));
    r.y = fe_sub(fp_mul(e, fe_sub(d, r.x)), c8);
    return r;
}

// P + Q, general Jacobian (12M + 4S)
HD G1J g1_add(const G1J& p, const G1J& q) {
    if (p.is_inf()) return q;
    if (q.is_inf()) return p;
    Fp z1z1 = fp_sqr(p.z), z2z2 = fp_sqr(q.z);
    Fp u1 = fp_mul(p.x, z2z2), u2 = fp_mul(q.x, z1z1);
    Fp s1 = fp_mul(fp_mul(p.y, q.z), z2z2), s2 = fp_mul(fp_mul(q.y, p.z), z1z1);
    if (u1 == u2) {
        if (s1 == s2) return g1_dbl(p);
        return G1J::infinity();
    }
    Fp h = fe_sub(u2, u1), rr = fe_sub(s2, s1);
    Fp hh = fp_sqr(h), hhh = fp_mul(h, hh), v = fp_mul(u1, hh);
    G1J r;
    r.x = fe_sub(fe_sub(fp_sqr(rr), hhh), fe_dbl(v));
    r.y = fe_sub(fp_mul(rr, fe_sub(v, r.x)), fp_mul(s1, hhh));
    r.z = fp_mul(fp_mul(p.z, q.z), h);
    return r;
}
HD G1J g1_sub(const G1J& p, const G1J& q) { return g1_add(p, g1_neg(q)); }

// P + Q with Q affine (8M + 3S)
HD G1J g1_add_mixed(const G1J& p, const G1A& q) {
    if (q.is_inf()) return p;
    if (p.is_inf()) { G1J r; r.x = q.x; r.y = q.y; r.z = Fp::one(); return r; }
    Fp z1z1 = fp_sqr(p.z);
    Fp u2 = fp_mul(q.x, z1z1), s2 = fp_mul(fp_mul(q.y, p.z), z1z1);
    if (p.x == u2) {
        if (p.y == s2) return g1_dbl(p);
        return G1J::infinity();
    }
    Fp h = fe_sub(u2, p.x), rr = fe_sub(s2, p.y);
    Fp hh = fp_sqr(h), hhh = fp_mul(h, hh), v = fp_mul(p.x, hh);
    G1J r;
    r.x = fe_sub(fe_sub(fp_sqr(rr), hhh), fe_dbl(v));
    r.y = fe_sub(fp_mul(rr, fe_sub(v, r.x)), fp_mul(p.y, hhh));
    r.z = fp_mul(p.z, h);
    return r;
}

// (x + t, x - t) sharing the common subexpressions of the two additions (18 mults instead
// of 32): the radix-2 butterfly of fft_g1.go:52-54.
HD void g1_add_sub(const G1J& x, const G1J& t, G1J& sum, G1J& diff) {
    if (t.is_inf()) { sum = x; diff = x; return; }
    if (x.is_inf()) { sum = t; diff = g1_neg(t); return; }
    Fp z1z1 = fp_sqr(x.z), z2z2 = fp_sqr(t.z);
    Fp u1 = fp_mul(x.x, z2z2), u2 = fp_mul(t.x, z1z1);
    Fp s1 = fp_mul(fp_mul(x.y, t.z), z2z2), s2 = fp_mul(fp_mul(t.y, x.z), z1z1);
    if (u1 == u2) {   // t == +-x: rare, take the generic path
        sum = g1_add(x, t);
        diff = g1_add(x, g1_neg(t));
        return;
    }
    Fp h = fe_sub(u2, u1);
    Fp hh = fp_sqr(h), hhh = fp_mul(h, hh), v = fp_mul(u1, hh);
    Fp z3 = fp_mul(fp_mul(x.z, t.z), h);
    Fp s1hhh = fp_mul(s1, hhh);
    Fp v2 = fe_dbl(v);
    Fp rp = fe_sub(s2, s1);                       // x + t
    Fp rm = fe_sub(fe_neg(s2), s1);               // x + (-t)
    sum.x = fe_sub(fe_sub(fp_sqr(rp), hhh), v2);
    sum.y = fe_sub(fp_mul(rp, fe_sub(v, sum.x)), s1hhh);
    sum.z = z3;
    diff.x = fe_sub(fe_sub(fp_sqr(rm), hhh), v2);
    diff.y = fe_sub(fp_mul(rm, fe_sub(v, diff.x)), s1hhh);
    diff.z = z3;
}

// endomorphism used by the GLV split k = k1 + k2 z^2:  z^2 (x, y) = (beta x, -y)
HD G1J g1_endo(const G1J& p) { G1J r; r.x = fp_mul(p.x, fp_const_beta()); r.y = fe_neg(p.y); r.z = p.z; return r; }

HD bool g1_equal(const G1J& a, const G1J& b) {   // bls/bls_kilic.go:106 EqualG1
    if (a.is_inf() || b.is_inf()) return a.is_inf() && b.is_inf();
    Fp za = fp_sqr(a.z), zb = fp_sqr(b.z);
    if (fp_mul(a.x, zb) != fp_mul(b.x, za)) return false;
    return fp_mul(a.y, fp_mul(zb, b.z)) == fp_mul(b.y, fp_mul(za, a.z));
}

// ---------------------------------------------------------------------------------------
// scalar multiplication
// ---------------------------------------------------------------------------------------
// k * P by plain double-and-add, k canonical 8 x u32 (host level-1 API: bls/bls_kilic.go:41 MulG1).
HD G1J g1_mul_simple(const G1J& p, const uint32_t* k) {
    G1J acc = G1J::infinity();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        if (started) acc = g1_dbl(acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) { acc = started ? g1_add(acc, p) : p; started = true; }
    }
    return acc;
}

// Prime-order subgroup membership.  phi(x, y) = (beta x, y) satisfies phi^2 + phi + 1 = 0 and acts on G1 as
// multiplication by -z^2 (z^4 - z^2 + 1 = r); the endomorphism phi + [z^2] has degree N(z^2 + omega) = r, so its
// kernel is exactly the r-torsion subgroup G1:  P in G1  <=>  [z^2] P == (beta x, -y).  A 128-bit multiplication
// instead of one by r.
HD bool g1_in_subgroup(const G1J& p) {
    if (p.is_inf()) return true;
    constexpr uint32_t z2l[4] = B200_GLV_Z2;
    uint32_t k[8] = {z2l[0], z2l[1], z2l[2], z2l[3], 0, 0, 0, 0};
    return g1_equal(g1_mul_simple(p, k), g1_endo(p));
}

// Scalar "program": a fixed scalar pre-split on the host as k = k1 + k2 z^2 (GLV) with both
// halves recoded into signed digits indexed by bit position (see g1_dev.cuh for the two modes).
#define B200_WNAF_LEN 132   // >= 130 digit positions per half scalar, padded to 4
struct ScalarProgram {
    int8_t d1[B200_WNAF_LEN];   // digits of k1
    int8_t d2[B200_WNAF_LEN];   // digits of k2 (applied to the endomorphism image)
    int16_t top;                // highest non-zero position (-1: scalar is zero)
    int8_t is_one;              // scalar == 1 (skip the multiplication)
    int8_t mode;                // 0 fixed 4-bit windows, 1 width-5 NAF
    int16_t pad[2];
};

}  // namespace b200
