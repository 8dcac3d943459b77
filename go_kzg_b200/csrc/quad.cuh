// quad.cuh -- point operations spread over the four lanes of a "quad" (lanes 4q .. 4q+3 of a warp).
//
// Why: a 381-bit Montgomery product is ~1900 cycles of dependent IMAD.WIDE carry chain for ONE warp, and a warp
// with one active lane pays the same as a full one (tools/lat_probe.cu: two independent products in one thread take
// exactly twice as long -- the multiplier pipe of the SM sub-partition is the limit, instruction-level parallelism
// inside a thread buys nothing).  The kernels that have few points to work on (tails of the bucket MSM, the
// transforms of a single polynomial) are bound by the DEPTH of the point formulas, not by throughput.  The formulas
// have width: a doubling is 7 products in 3 dependent levels, an addition 16 in 5.  Here the lanes of a quad each
// take one product of a level; the results are all-gathered with shuffles (48 per level, noise against a product),
// the cheap linear steps run redundantly in all four lanes.  The state of a point is replicated across its quad.
// A quad operation costs (levels x one product) instead of (products x one product): doubling 3 instead of 7,
// addition 5 instead of 16, mixed addition 5 instead of 11.
//
// Every operation here is WARP-COLLECTIVE: all 32 lanes call it together and the shuffles carry the constant full
// mask (a run-time mask makes nvcc wrap every shuffle in MATCH / VOTE / WARPSYNC, ~40 cycles each, which costs more
// than the products saved -- measured).  Quads that have nothing to do pass active = false and keep their value;
// the special cases of the group law are handled by selection instead of branches (an addition that turns out to be
// a doubling makes the whole warp run one extra doubling, which is rare).
#pragma once
#include "g1_dev.cuh"

namespace b200 {

#define QUAD_FULL 0xffffffffu
__device__ __forceinline__ unsigned quad_role() { return threadIdx.x & 3u; }

// Operand selection by lane role WITHOUT branches: ternaries over 12-word structs that live in local memory compile to
// divergent branches (BSSY / BRA / BSYNC around the loads; the first form of this file had 98 such regions in one point
// addition, each serialising the four lanes of every quad).  Masks and LOP3 instead.
__device__ __forceinline__ Fp quad_pick2(unsigned sel, const Fp& v0, const Fp& v1) {            // sel ? v1 : v0
    const uint32_t m = 0u - (sel & 1u);
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = (v0.l[i] & ~m) | (v1.l[i] & m);
    return r;
}
__device__ __forceinline__ Fp quad_pick3(unsigned role, const Fp& v0, const Fp& v1, const Fp& v2) {   // role 3 -> v2 as well
    const uint32_t m1 = 0u - (uint32_t)(role == 1), m2 = 0u - (uint32_t)(role >= 2);
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = (v0.l[i] & ~(m1 | m2)) | (v1.l[i] & m1) | (v2.l[i] & m2);
    return r;
}
__device__ __forceinline__ Fp quad_pick(unsigned role, const Fp& v0, const Fp& v1, const Fp& v2, const Fp& v3) {
    const uint32_t lo = 0u - (role & 1u), hi = 0u - ((role >> 1) & 1u);
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        const uint32_t a = (v0.l[i] & ~lo) | (v1.l[i] & lo), b = (v2.l[i] & ~lo) | (v3.l[i] & lo);
        r.l[i] = (a & ~hi) | (b & hi);
    }
    return r;
}
__device__ __forceinline__ Fp quad_bcast(const Fp& v, int src) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(QUAD_FULL, v.l[i], src, 4);
    return r;
}
__device__ __forceinline__ G1J g1_select(bool c, const G1J& a, const G1J& b) {   // c ? a : b, word by word, no branch
    const uint32_t m = 0u - (uint32_t)c;
    G1J r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r.x.l[i] = (a.x.l[i] & m) | (b.x.l[i] & ~m);
        r.y.l[i] = (a.y.l[i] & m) | (b.y.l[i] & ~m);
        r.z.l[i] = (a.z.l[i] & m) | (b.z.l[i] & ~m);
    }
    return r;
}
// one level of K distinct products: lane j of the quad computes product min(j, K - 1) (K = 2: lanes 0, 2 the first and 1, 3
// the second), only the K results are gathered (12 shuffles each).  A level whose lanes would all compute the SAME
// product is a plain fp_mul in every lane: no selection, no shuffles.
__device__ __forceinline__ void quad_mul4(const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, const Fp& a3,
                                          const Fp& b3, Fp& r0, Fp& r1, Fp& r2, Fp& r3) {
    const unsigned role = quad_role();
    const Fp p = fp_mul(quad_pick(role, a0, a1, a2, a3), quad_pick(role, b0, b1, b2, b3));
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1); r2 = quad_bcast(p, 2); r3 = quad_bcast(p, 3);
}
__device__ __forceinline__ void quad_mul3(const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, Fp& r0, Fp& r1,
                                          Fp& r2) {
    const unsigned role = quad_role();
    const Fp p = fp_mul(quad_pick3(role, a0, a1, a2), quad_pick3(role, b0, b1, b2));
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1); r2 = quad_bcast(p, 2);
}
__device__ __forceinline__ void quad_mul2(const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, Fp& r0, Fp& r1) {
    const unsigned role = quad_role();
    const Fp p = fp_mul(quad_pick2(role, a0, a1), quad_pick2(role, b0, b1));
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1);
}
__device__ __forceinline__ void quad_sqr3(const Fp& a0, const Fp& a1, const Fp& a2, Fp& r0, Fp& r1, Fp& r2) {
    const Fp p = fp_sqr(quad_pick3(quad_role(), a0, a1, a2));
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1); r2 = quad_bcast(p, 2);
}

// *p = 2 * *p where active, in 3 levels (formulas of g1_dbl: 2M + 5S; they map infinity to infinity by themselves)
static __device__ __noinline__ void quad_dbl(G1J* p_io, bool active) {
    const G1J p = *p_io;
    Fp a, b, yz, c, t, f;
    quad_mul3(p.x, p.x, p.y, p.y, p.y, p.z, a, b, yz);
    const Fp e = fe_add(fe_dbl(a), a);
    quad_sqr3(b, fe_add(p.x, b), e, c, t, f);
    const Fp d = fe_dbl(fe_sub(fe_sub(t, a), c));
    G1J o;
    o.z = fe_dbl(yz);
    o.x = fe_sub(f, fe_dbl(d));
    const Fp ew = fp_mul(e, fe_sub(d, o.x));                  // the same product in every lane
    o.y = fe_sub(ew, fe_dbl(fe_dbl(fe_dbl(c))));
    *p_io = g1_select(active, o, p);
}

// *p = *p + *q where active; both Jacobian; 5 levels.  Cases of g1_add by selection: an infinite operand returns the
// other one; P == -Q falls out of the formulas (H = 0 gives Z3 = 0); P == Q needs a doubling, run by the whole warp
// when any active quad hits it.
static __device__ __noinline__ void quad_add(G1J* p_io, const G1J* q_in, bool active) {
    const G1J p = *p_io, q = *q_in;
    Fp z1z1, z2z2, a, b, u1, u2, s1, s2;
    quad_mul4(p.z, p.z, q.z, q.z, p.y, q.z, q.y, p.z, z1z1, z2z2, a, b);
    quad_mul4(p.x, z2z2, q.x, z1z1, a, z2z2, b, z1z1, u1, u2, s1, s2);
    const Fp h = fe_sub(u2, u1), rr = fe_sub(s2, s1);
    Fp hh, zz, r2, t0, hhh, v, z3, t1;
    quad_mul3(h, h, p.z, q.z, rr, rr, hh, zz, r2);
    quad_mul3(h, hh, u1, hh, zz, h, hhh, v, z3);
    G1J o;
    o.x = fe_sub(fe_sub(r2, hhh), fe_dbl(v));
    const Fp w = fe_sub(v, o.x);
    quad_mul2(rr, w, s1, hhh, t0, t1);
    o.y = fe_sub(t0, t1);
    o.z = z3;
    const bool pinf = p.is_inf(), qinf = q.is_inf();
    const bool same = active && !pinf && !qinf && h.is_zero() && rr.is_zero();
    if (__any_sync(QUAD_FULL, same)) {
        G1J d = p;
        quad_dbl(&d, true);
        o = g1_select(same, d, o);
    }
    o = g1_select(qinf, p, o);
    o = g1_select(pinf, q, o);
    *p_io = g1_select(active, o, p);
}

// *p = *p + *q where active, Q affine (finite, or (0, 0) for infinity); 5 levels
static __device__ __noinline__ void quad_add_mixed(G1J* p_io, const G1A* q_in, bool active) {
    const G1J p = *p_io;
    const G1A q = *q_in;
    Fp z1z1, t, u2, s2, t0, t1;
    quad_mul2(p.z, p.z, q.y, p.z, z1z1, t);
    quad_mul2(q.x, z1z1, t, z1z1, u2, s2);
    const Fp h = fe_sub(u2, p.x), rr = fe_sub(s2, p.y);
    Fp hh, r2, z3, hhh, v;
    quad_mul3(h, h, rr, rr, p.z, h, hh, r2, z3);
    quad_mul2(h, hh, p.x, hh, hhh, v);
    G1J o;
    o.x = fe_sub(fe_sub(r2, hhh), fe_dbl(v));
    const Fp w = fe_sub(v, o.x);
    quad_mul2(rr, w, p.y, hhh, t0, t1);
    o.y = fe_sub(t0, t1);
    o.z = z3;
    const bool pinf = p.is_inf(), qinf = q.is_inf();
    const bool same = active && !pinf && !qinf && h.is_zero() && rr.is_zero();
    if (__any_sync(QUAD_FULL, same)) {
        G1J d = p;
        quad_dbl(&d, true);
        o = g1_select(same, d, o);
    }
    G1J qj; qj.x = q.x; qj.y = q.y; qj.z = Fp::one();
    o = g1_select(qinf, p, o);
    o = g1_select(pinf && !qinf, qj, o);
    *p_io = g1_select(active, o, p);
}

// *out = m * *p where active, for a small m (at most 16 bits); warp-collective like everything here
__device__ __forceinline__ void quad_small_mul(G1J* out, const G1J* p, unsigned m, bool active) {
    G1J acc = G1J::infinity();
    const unsigned any_m = __reduce_or_sync(QUAD_FULL, active ? m : 0u);
    for (int bit = 31 - __clz(any_m | 1u); bit >= 0; bit--) {
        quad_dbl(&acc, active);
        quad_add(&acc, p, active && ((m >> bit) & 1u));
    }
    if (active) *out = acc;
}

// (x + t, x - t) in 5 levels, sharing the common subexpressions of the two additions (the radix-2 butterfly of
// fft_g1.go:52-54; g1_add_sub).  t == +-x is rare: the whole warp then runs two plain additions.
static __device__ __noinline__ void quad_add_sub(G1J* sum, G1J* diff, const G1J* x_in, const G1J* t_in) {
    const G1J x = *x_in, t = *t_in;
    Fp z1z1, z2z2, a, b, u1, u2, s1, s2;
    quad_mul4(x.z, x.z, t.z, t.z, x.y, t.z, t.y, x.z, z1z1, z2z2, a, b);
    quad_mul4(x.x, z2z2, t.x, z1z1, a, z2z2, b, z1z1, u1, u2, s1, s2);
    const Fp h = fe_sub(u2, u1), rp = fe_sub(s2, s1), rm = fe_sub(fe_neg(s2), s1);
    Fp hh, zz, rp2, rm2, hhh, v, z3;
    quad_mul4(h, h, x.z, t.z, rp, rp, rm, rm, hh, zz, rp2, rm2);
    quad_mul3(h, hh, u1, hh, zz, h, hhh, v, z3);
    const Fp v2 = fe_dbl(v);
    G1J sp, dm;
    sp.x = fe_sub(fe_sub(rp2, hhh), v2);
    dm.x = fe_sub(fe_sub(rm2, hhh), v2);
    Fp yp, ym, sh;
    quad_mul3(rp, fe_sub(v, sp.x), rm, fe_sub(v, dm.x), s1, hhh, yp, ym, sh);
    sp.y = fe_sub(yp, sh); dm.y = fe_sub(ym, sh);
    sp.z = z3; dm.z = z3;
    const bool xinf = x.is_inf(), tinf = t.is_inf();
    const bool same = !xinf && !tinf && h.is_zero();
    if (__any_sync(QUAD_FULL, same)) {
        G1J s2p = x, d2p = x, nt = t;
        nt.y = fe_neg(t.y);
        quad_add(&s2p, &t, true);
        quad_add(&d2p, &nt, true);
        sp = g1_select(same, s2p, sp);
        dm = g1_select(same, d2p, dm);
    }
    G1J nt = t;
    nt.y = fe_neg(t.y);
    sp = g1_select(xinf, t, sp); dm = g1_select(xinf, nt, dm);        // 0 + t, 0 - t
    sp = g1_select(tinf, x, sp); dm = g1_select(tinf, x, dm);         // x +- 0
    *sum = sp; *diff = dm;
}

// k * P for the digit strings of a ScalarProgram (g1_dev.cuh), one quad per product.  Jacobian table: with quads a general
// addition costs the same 5 levels as a mixed one, so the effective-affine construction of g1_mul_digits buys nothing here.
// mode 1 (sparse digits) requires every quad of the warp to hold the SAME program (across-block lane mapping); mode 0 adds at
// the fixed window positions for whichever quads have a non-zero digit there.
static __device__ __noinline__ void quad_mul_digits(G1J* out, const G1J* p_in, const int8_t* d1, const int8_t* d2, int top, int mode) {
    G1J tab[8];
    Fp bx[8];
    tab[0] = *p_in;
    if (mode == 0) {                      // {1..8} P
        tab[1] = tab[0]; quad_dbl(&tab[1], true);
        tab[2] = tab[1]; quad_add(&tab[2], &tab[0], true);
        tab[3] = tab[1]; quad_dbl(&tab[3], true);
        tab[4] = tab[3]; quad_add(&tab[4], &tab[0], true);
        tab[5] = tab[2]; quad_dbl(&tab[5], true);
        tab[6] = tab[5]; quad_add(&tab[6], &tab[0], true);
        tab[7] = tab[3]; quad_dbl(&tab[7], true);
    } else {                              // {1, 3, .., 15} P
        G1J p2 = tab[0];
        quad_dbl(&p2, true);
        for (int i = 1; i < 8; i++) { tab[i] = tab[i - 1]; quad_add(&tab[i], &p2, true); }
    }
    const Fp beta = fp_const_beta();
    quad_mul4(tab[0].x, beta, tab[1].x, beta, tab[2].x, beta, tab[3].x, beta, bx[0], bx[1], bx[2], bx[3]);
    quad_mul4(tab[4].x, beta, tab[5].x, beta, tab[6].x, beta, tab[7].x, beta, bx[4], bx[5], bx[6], bx[7]);
    G1J acc = G1J::infinity();
    const int wtop = __reduce_max_sync(QUAD_FULL, top);
    for (int i = wtop; i >= 0; i--) {
        quad_dbl(&acc, true);             // the doubling formulas keep infinity at infinity
        const int a = d1[i];
        if (__any_sync(QUAD_FULL, a != 0)) {
            const int mg = a < 0 ? -a : a;
            const int idx = a ? (mode == 0 ? mg - 1 : mg >> 1) : 0;
            G1J t = tab[idx];
            if (a < 0) t.y = fe_neg(t.y);
            quad_add(&acc, &t, a != 0);
        }
        const int b = d2[i];
        if (__any_sync(QUAD_FULL, b != 0)) {
            const int mg = b < 0 ? -b : b;
            const int idx = b ? (mode == 0 ? mg - 1 : mg >> 1) : 0;
            G1J t = tab[idx];
            t.x = bx[idx];
            if (b > 0) t.y = fe_neg(t.y);     // z^2 (x, y) = (beta x, -y)
            quad_add(&acc, &t, b != 0);
        }
    }
    *out = acc;
}
// warp-collective g1_mul_program: a product by 1 is skipped only when the whole warp has it
__device__ __forceinline__ void quad_mul_program(G1J* out, const G1J* p, const ScalarProgram* prog) {
    if (__all_sync(QUAD_FULL, prog->is_one != 0)) { *out = *p; return; }
    G1J r;
    quad_mul_digits(&r, p, prog->d1, prog->d2, prog->top, prog->mode);
    *out = r;
}

// sum over the quads of a warp: every quad ends with the total of all eight
__device__ __forceinline__ G1J quad_shfl_xor(const G1J& p, unsigned lane_mask) {
    G1J r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r.x.l[i] = __shfl_xor_sync(QUAD_FULL, p.x.l[i], lane_mask);
        r.y.l[i] = __shfl_xor_sync(QUAD_FULL, p.y.l[i], lane_mask);
        r.z.l[i] = __shfl_xor_sync(QUAD_FULL, p.z.l[i], lane_mask);
    }
    return r;
}
// groups of `quads` adjacent quads (power of two <= 8)
__device__ __forceinline__ void quad_group_sum(G1J& acc, unsigned quads) {
    for (unsigned off = (quads * 4) >> 1; off >= 4; off >>= 1) {
        G1J other = quad_shfl_xor(acc, off);
        quad_add(&acc, &other, true);
    }
}
__device__ __forceinline__ void quad_warp_sum(G1J& acc) { quad_group_sum(acc, 8); }

}  // namespace b200
