// quad.cuh -- point operations spread over the four lanes of a "quad" (lanes 4q .. 4q+3 of a warp).
//
// Why: a 381-bit Montgomery product is ~1900 cycles of dependent IMAD.WIDE carry chain for ONE warp, and a warp
// with one active lane pays the same as a full one (tools/lat_probe.cu: two independent products in one thread take
// exactly twice as long -- the multiplier pipe of the SM sub-partition is the limit, instruction-level parallelism
// inside a thread buys nothing).  The kernels that have few points to work on (tails of the bucket MSM, the
// transforms of a single polynomial) are bound by the DEPTH of the point formulas, not by throughput.  The formulas
// have width: a doubling is 7 products in 3 dependent levels, an addition 16 in 5.  Here the lanes of a quad each
// take one product of a level; the results are all-gathered with shuffles (48 per level, noise against a product),
// the cheap linear steps run redundantly in all four lanes.  The state of a point is replicated across its quad,
// so control flow is uniform inside a quad.  A quad operation costs (levels x one product) instead of
// (products x one product): doubling 3 instead of 7, addition 5 instead of 16, mixed addition 5 instead of 11.
//
// The four lanes of a quad must call together (blocks are one-dimensional with a multiple of 32 threads); different
// quads are independent.
#pragma once
#include "g1_dev.cuh"

namespace b200 {

__device__ __forceinline__ unsigned quad_role() { return threadIdx.x & 3u; }

__device__ __forceinline__ Fp quad_pick(unsigned role, const Fp& v0, const Fp& v1, const Fp& v2, const Fp& v3) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = role == 0 ? v0.l[i] : (role == 1 ? v1.l[i] : (role == 2 ? v2.l[i] : v3.l[i]));
    return r;
}
// The shuffles of a quad name only its own four lanes: quads of one warp may sit in different branches of the
// group law (infinity, doubling) without waiting for each other.
__device__ __forceinline__ unsigned quad_mask() { return 0xFu << (threadIdx.x & 28u); }
__device__ __forceinline__ Fp quad_bcast(const Fp& v, int src) {
    const unsigned m = quad_mask();
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(m, v.l[i], src, 4);
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

// 2P in 3 levels (same formulas as g1_dbl: 2M + 5S)
__device__ __noinline__ void quad_dbl(G1J* r, const G1J* p_in) {
    const G1J p = *p_in;
    if (p.is_inf()) { *r = p; return; }
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
    *r = o;
}

// P + Q, both Jacobian, in 5 levels (same case analysis as g1_add)
__device__ __noinline__ void quad_add(G1J* r, const G1J* p_in, const G1J* q_in) {
    const G1J p = *p_in, q = *q_in;
    if (p.is_inf()) { *r = q; return; }
    if (q.is_inf()) { *r = p; return; }
    Fp z1z1, z2z2, a, b, u1, u2, s1, s2;
    quad_mul(p.z, p.z, q.z, q.z, p.y, q.z, q.y, p.z, z1z1, z2z2, a, b);
    quad_mul(p.x, z2z2, q.x, z1z1, a, z2z2, b, z1z1, u1, u2, s1, s2);
    if (u1 == u2) {
        if (s1 == s2) { quad_dbl(r, p_in); return; }
        *r = G1J::infinity();
        return;
    }
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
    *r = o;
}

// P + Q with Q affine (finite or the (0, 0) infinity), in 5 levels (same case analysis as g1_add_mixed)
__device__ __noinline__ void quad_add_mixed(G1J* r, const G1J* p_in, const G1A* q_in) {
    const G1J p = *p_in;
    const G1A q = *q_in;
    if (q.is_inf()) { *r = p; return; }
    if (p.is_inf()) { G1J o; o.x = q.x; o.y = q.y; o.z = Fp::one(); *r = o; return; }
    Fp z1z1, t, u2, s2, t0, t1;
    quad_mul(p.z, p.z, q.y, p.z, p.z, p.z, q.y, p.z, z1z1, t, t0, t1);
    quad_mul(q.x, z1z1, t, z1z1, q.x, z1z1, t, z1z1, u2, s2, t0, t1);
    if (p.x == u2) {
        if (p.y == s2) { quad_dbl(r, p_in); return; }
        *r = G1J::infinity();
        return;
    }
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
    *r = o;
}

// m * P for a small m (at most 16 bits)
__device__ __forceinline__ void quad_small_mul(G1J* out, const G1J* p, unsigned m) {
    G1J acc = G1J::infinity();
    for (int bit = 15; bit >= 0; bit--) {
        if (!acc.is_inf()) quad_dbl(&acc, &acc);
        if ((m >> bit) & 1u) quad_add(&acc, &acc, p);
    }
    *out = acc;
}

// sum over the quads of a warp: every quad ends with the total of all eight (quad-replicated values stay replicated)
__device__ __forceinline__ G1J quad_shfl_xor(const G1J& p, unsigned lane_mask) {
    G1J r;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        r.x.l[i] = __shfl_xor_sync(0xffffffffu, p.x.l[i], lane_mask);
        r.y.l[i] = __shfl_xor_sync(0xffffffffu, p.y.l[i], lane_mask);
        r.z.l[i] = __shfl_xor_sync(0xffffffffu, p.z.l[i], lane_mask);
    }
    return r;
}
__device__ __forceinline__ void quad_warp_sum(G1J& acc) {
    for (unsigned off = 16; off >= 4; off >>= 1) {
        G1J other = quad_shfl_xor(acc, off);
        quad_add(&acc, &acc, &other);
    }
}

}  // namespace b200
