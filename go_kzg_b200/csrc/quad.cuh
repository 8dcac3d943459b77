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

__device__ __forceinline__ Fp quad_pick(unsigned role, const Fp& v0, const Fp& v1, const Fp& v2, const Fp& v3) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = role == 0 ? v0.l[i] : (role == 1 ? v1.l[i] : (role == 2 ? v2.l[i] : v3.l[i]));
    return r;
}
__device__ __forceinline__ Fp quad_bcast(const Fp& v, int src) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(QUAD_FULL, v.l[i], src, 4);
    return r;
}
__device__ __forceinline__ G1J g1_select(bool c, const G1J& a, const G1J& b) {   // c ? a : b, word by word
    G1J r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r.x.l[i] = c ? a.x.l[i] : b.x.l[i];
        r.y.l[i] = c ? a.y.l[i] : b.y.l[i];
        r.z.l[i] = c ? a.z.l[i] : b.z.l[i];
    }
    return r;
}
// one level: r_j = a_j * b_j for j < 4, lane j of the quad computing product j (unused slots: pass any operands)
__device__ __forceinline__ void quad_mul(const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2, const Fp& a3,
                                         const Fp& b3, Fp& r0, Fp& r1, Fp& r2, Fp& r3) {
    const unsigned role = quad_role();
    const Fp a = quad_pick(role, a0, a1, a2, a3), b = quad_pick(role, b0, b1, b2, b3);
    const Fp p = fp_mul(a, b);
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1); r2 = quad_bcast(p, 2); r3 = quad_bcast(p, 3);
}
// one level of squares
__device__ __forceinline__ void quad_sqr(const Fp& a0, const Fp& a1, const Fp& a2, const Fp& a3, Fp& r0, Fp& r1, Fp& r2, Fp& r3) {
    const Fp p = fp_sqr(quad_pick(quad_role(), a0, a1, a2, a3));
    r0 = quad_bcast(p, 0); r1 = quad_bcast(p, 1); r2 = quad_bcast(p, 2); r3 = quad_bcast(p, 3);
}

// *p = 2 * *p where active, in 3 levels (formulas of g1_dbl: 2M + 5S; they map infinity to infinity by themselves)
__device__ __noinline__ void quad_dbl(G1J* p_io, bool active) {
    const G1J p = *p_io;
    Fp a, b, yz, u0, c, t, f;
    quad_mul(p.x, p.x, p.y, p.y, p.y, p.z, p.x, p.x, a, b, yz, u0);
    const Fp e = fe_add(fe_dbl(a), a);
    quad_sqr(b, fe_add(p.x, b), e, e, c, t, f, u0);
    const Fp d = fe_dbl(fe_sub(fe_sub(t, a), c));
    G1J o;
    o.z = fe_dbl(yz);
    o.x = fe_sub(f, fe_dbl(d));
    const Fp w = fe_sub(d, o.x);
    Fp ew, u1, u2;
    quad_mul(e, w, e, w, e, w, e, w, ew, u0, u1, u2);
    o.y = fe_sub(ew, fe_dbl(fe_dbl(fe_dbl(c))));
    if (active) *p_io = o;
}

// *p = *p + *q where active; both Jacobian; 5 levels.  Cases of g1_add by selection: an infinite operand returns the
// other one; P == -Q falls out of the formulas (H = 0 gives Z3 = 0); P == Q needs a doubling, run by the whole warp
// when any active quad hits it.
__device__ __noinline__ void quad_add(G1J* p_io, const G1J* q_in, bool active) {
    const G1J p = *p_io, q = *q_in;
    Fp z1z1, z2z2, a, b, u1, u2, s1, s2;
    quad_mul(p.z, p.z, q.z, q.z, p.y, q.z, q.y, p.z, z1z1, z2z2, a, b);
    quad_mul(p.x, z2z2, q.x, z1z1, a, z2z2, b, z1z1, u1, u2, s1, s2);
    const Fp h = fe_sub(u2, u1), rr = fe_sub(s2, s1);
    Fp hh, zz, r2, t0, hhh, v, z3, t1;
    quad_mul(h, h, p.z, q.z, rr, rr, h, h, hh, zz, r2, t0);
    quad_mul(h, hh, u1, hh, zz, h, h, hh, hhh, v, z3, t0);
    G1J o;
    o.x = fe_sub(fe_sub(r2, hhh), fe_dbl(v));
    const Fp w = fe_sub(v, o.x);
    quad_mul(rr, w, s1, hhh, rr, w, s1, hhh, t0, t1, hh, zz);
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
    if (active) *p_io = o;
}

// *p = *p + *q where active, Q affine (finite, or (0, 0) for infinity); 5 levels
__device__ __noinline__ void quad_add_mixed(G1J* p_io, const G1A* q_in, bool active) {
    const G1J p = *p_io;
    const G1A q = *q_in;
    Fp z1z1, t, u2, s2, t0, t1;
    quad_mul(p.z, p.z, q.y, p.z, p.z, p.z, q.y, p.z, z1z1, t, t0, t1);
    quad_mul(q.x, z1z1, t, z1z1, q.x, z1z1, t, z1z1, u2, s2, t0, t1);
    const Fp h = fe_sub(u2, p.x), rr = fe_sub(s2, p.y);
    Fp hh, r2, z3, hhh, v;
    quad_mul(h, h, rr, rr, p.z, h, h, h, hh, r2, z3, t0);
    quad_mul(h, hh, p.x, hh, h, hh, p.x, hh, hhh, v, t0, t1);
    G1J o;
    o.x = fe_sub(fe_sub(r2, hhh), fe_dbl(v));
    const Fp w = fe_sub(v, o.x);
    quad_mul(rr, w, p.y, hhh, rr, w, p.y, hhh, t0, t1, hh, r2);
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
    if (active) *p_io = o;
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
