// g1_dev.cuh -- device-only G1 building blocks shared by the kernels: out-of-line point
// operations (one copy of the ~10k-instruction add / double bodies per kernel image instead of
// one per call site), digit-programmed scalar multiplication, and on-device GLV recoding of
// variable scalars.
//
// Replaces (semantics): bls.MulG1 / AddG1 / SubG1 as used by fft_g1.go:47-55 (butterfly),
// fk20_single.go:72-74 (ToeplitzPart2) and bls/bls_kilic.go:132-150 (LinCombG1) of the reference.
#pragma once
#include "g1.cuh"

namespace b200 {

// ---- out-of-line group law ---------------------------------------------------------------
static __host__ __device__ __noinline__ void g1_dbl_ni(G1J* r, const G1J* p) { *r = g1_dbl(*p); }
static __host__ __device__ __noinline__ void g1_add_ni(G1J* r, const G1J* p, const G1J* q) { *r = g1_add(*p, *q); }
static __host__ __device__ __noinline__ void g1_add_mixed_ni(G1J* r, const G1J* p, const G1A* q) { *r = g1_add_mixed(*p, *q); }
// (x + t, x - t): the radix-2 butterfly of fft_g1.go:52-54, sharing the common subexpressions of
// the two additions (18 products instead of 32).
static __host__ __device__ __noinline__ void g1_add_sub_ni(G1J* sum, G1J* diff, const G1J* x, const G1J* t) {
    G1J s, d;
    g1_add_sub(*x, *t, s, d);
    *sum = s; *diff = d;
}

// ---- digit-programmed scalar multiplication ----------------------------------------------
// A scalar k = k1 + k2 z^2 (GLV, both halves < 2^128) is given as two signed digit strings
// indexed by bit position: k_h = sum_i d_h[i] 2^i.  Two recodings share this routine:
//   mode 0  fixed 4-bit windows, digits in [-8, 8] at positions 0,4,8,..: every lane of a warp
//           adds at the same positions whatever its scalar -> used when lanes hold different
//           scalars (variable scalars, single-transform G1 FFT).     table = {1..8} P
//   mode 1  width-5 NAF, odd digits in [-15, 15]: fewest additions, used when the whole warp
//           shares one fixed scalar (batched G1 FFT: lanes = blobs).  table = {1,3,..,15} P
// z^2 (x, y) = (beta x, -y).
//
// Reference path (Jacobian table, general additions).  Only taken when the table construction
// below meets a degenerate addition, i.e. for points outside the prime-order subgroup.
// Out of line on purpose: with this body inlined next to the call that produced *p, nvcc 12.9
// emitted the copy `tab[0] = *p` without its first 16 bytes (wrong twiddle products).
static __host__ __device__ __noinline__ void g1_mul_digits_jac(G1J* out, const G1J* p, const int8_t* d1, const int8_t* d2, int top,
                                                                int mode) {
    if (top < 0 || p->is_inf()) { *out = G1J::infinity(); return; }
    G1J tab[8];
    Fp bx[8];
    tab[0] = *p;
    if (mode == 0) {
        g1_dbl_ni(&tab[1], &tab[0]);
        g1_add_ni(&tab[2], &tab[1], &tab[0]);
        g1_dbl_ni(&tab[3], &tab[1]);
        g1_add_ni(&tab[4], &tab[3], &tab[0]);
        g1_dbl_ni(&tab[5], &tab[2]);
        g1_add_ni(&tab[6], &tab[5], &tab[0]);
        g1_dbl_ni(&tab[7], &tab[3]);
    } else {
        G1J p2;
        g1_dbl_ni(&p2, &tab[0]);
        for (int i = 1; i < 8; i++) g1_add_ni(&tab[i], &tab[i - 1], &p2);
    }
    const Fp beta = fp_const_beta();
    for (int i = 0; i < 8; i++) bx[i] = fp_mul(tab[i].x, beta);
    G1J acc = G1J::infinity();
    G1J t;
    for (int i = top; i >= 0; i--) {
        if (!acc.is_inf()) g1_dbl_ni(&acc, &acc);
        int a = d1[i];
        if (a) {
            int m = a < 0 ? -a : a;
            int idx = mode == 0 ? m - 1 : m >> 1;
            t = tab[idx];
            if (a < 0) t.y = fe_neg(t.y);
            g1_add_ni(&acc, &acc, &t);
        }
        int b = d2[i];
        if (b) {
            int m = b < 0 ? -b : b;
            int idx = mode == 0 ? m - 1 : m >> 1;
            t = tab[idx];
            t.x = bx[idx];
            if (b > 0) t.y = fe_neg(t.y);
            g1_add_ni(&acc, &acc, &t);
        }
    }
    *out = acc;
}

// r = p + q (q affine) for finite p, q with p != +-q, also handing back H (Z_r = Z_p * H).
// Returns false (r untouched) in the degenerate cases.
static __host__ __device__ __noinline__ bool g1_add_mixed_h(G1J* r, Fp* h_out, const G1J* p, const G1A* q) {
    Fp z1z1 = fp_sqr(p->z);
    Fp u2 = fp_mul(q->x, z1z1), s2 = fp_mul(fp_mul(q->y, p->z), z1z1);
    Fp h = fe_sub(u2, p->x);
    if (h.is_zero() || p->z.is_zero()) return false;
    Fp rr = fe_sub(s2, p->y);
    Fp hh = fp_sqr(h), hhh = fp_mul(h, hh), v = fp_mul(p->x, hh);
    G1J o;
    o.x = fe_sub(fe_sub(fp_sqr(rr), hhh), fe_dbl(v));
    o.y = fe_sub(fp_mul(rr, fe_sub(v, o.x)), fp_mul(p->y, hhh));
    o.z = fp_mul(p->z, h);
    *r = o;
    *h_out = h;
    return true;
}

// The hot routine.  The look-up table is built and used on an isomorphic curve on which all of
// its entries are AFFINE, so that every addition of the main loop is a mixed addition (8M + 3S
// instead of 12M + 4S; ~43 of them per product):
//   * y^2 = x^3 + b and y^2 = x^3 + b u^6 are isomorphic through (x, y) -> (u^2 x, u^3 y), and the
//     Jacobian doubling / addition formulas of an a = 0 curve do not contain b.  A point with
//     Jacobian coordinates (X, Y, Z) on the u-curve is (X, Y, Z u) on the original one.
//   * the chain tab[i] = tab[i-1] + S is run with S affine on the curve scaled by Z_S (mixed
//     additions), every step reporting the ratio H_i = Z_i / Z_(i-1); walking back, entry i is
//     rescaled by (Z_7 / Z_i)^(2,3), which makes all entries affine on the curve scaled by
//     Zg = Z_7 * Z_S.  The endomorphism (x, y) -> (beta x, -y) commutes with the scaling.
//   * the accumulator lives on that curve; one product by Zg brings the result home.
// ext (optional): storage for the 16 coordinates of the table outside the thread's stack -- shared memory in the
// B200_STAGE_SMEM_TABLE build of the stage kernel; coordinate c of entry i lives at ext[(c * 8 + i) * estride].
static __host__ __device__ __noinline__ void g1_mul_digits(G1J* out, const G1J* p, const int8_t* d1, const int8_t* d2, int top,
                                                            int mode, Fp* ext = nullptr, unsigned estride = 1) {
    if (top < 0 || p->is_inf()) { *out = G1J::infinity(); return; }
    Fp xy_local[16];  // table entries x[0..8), y[0..8) (untouched when ext is given)
    Fp hz[8];         // the ratio H_i while the table is being built
    Fp bx[8];
    Fp* const T = ext ? ext : xy_local;
    const unsigned ts = ext ? estride : 1u;
#define TAB_X(i) T[(unsigned)(i) * ts]
#define TAB_Y(i) T[(8u + (unsigned)(i)) * ts]
    G1J cur, nxt;
    G1A step;
    Fp zs;            // Z_S: scaling of the curve the chain runs on
    int first;
    if (mode == 0) {  // {1..8} P: S = P, chain starts at 2P
        zs = p->z;
        step.x = p->x; step.y = p->y;
        cur.x = p->x; cur.y = p->y; cur.z = Fp::one();
        TAB_X(0) = cur.x; TAB_Y(0) = cur.y; hz[0] = cur.z;
        g1_dbl_ni(&nxt, &cur);
        TAB_X(1) = nxt.x; TAB_Y(1) = nxt.y; hz[1] = nxt.z;   // z = Z_1 / Z_0 with Z_0 = 1
        cur = nxt;
        first = 2;
    } else {          // {1,3,..,15} P: S = 2P, chain starts at P
        g1_dbl_ni(&nxt, p);
        zs = nxt.z;
        step.x = nxt.x; step.y = nxt.y;
        Fp c2 = fp_sqr(zs);
        cur.x = fp_mul(p->x, c2); cur.y = fp_mul(p->y, fp_mul(c2, zs)); cur.z = p->z;
        TAB_X(0) = cur.x; TAB_Y(0) = cur.y; hz[0] = cur.z;
        first = 1;
    }
    bool ok = !zs.is_zero();
    for (int i = first; i < 8 && ok; i++) {
        Fp h;
        ok = g1_add_mixed_h(&nxt, &h, &cur, &step);
        TAB_X(i) = nxt.x; TAB_Y(i) = nxt.y; hz[i] = h;
        cur = nxt;
    }
    if (!ok) { g1_mul_digits_jac(out, p, d1, d2, top, mode); return; }
    const Fp zg = fp_mul(cur.z, zs);
    {
        Fp s = hz[7];                      // Z_7 / Z_6
        for (int i = 6; i >= 0; i--) {
            Fp s2 = fp_sqr(s);
            TAB_X(i) = fp_mul(TAB_X(i), s2);
            TAB_Y(i) = fp_mul(TAB_Y(i), fp_mul(s2, s));
            if (i) s = fp_mul(s, hz[i]);
        }
    }
    const Fp beta = fp_const_beta();
    for (int i = 0; i < 8; i++) bx[i] = fp_mul(TAB_X(i), beta);
    G1J acc = G1J::infinity();
    G1A t;
    for (int i = top; i >= 0; i--) {
        if (!acc.is_inf()) g1_dbl_ni(&acc, &acc);
        int a = d1[i];
        if (a) {
            int m = a < 0 ? -a : a;
            int idx = mode == 0 ? m - 1 : m >> 1;
            t.x = TAB_X(idx); t.y = TAB_Y(idx);
            if (a < 0) t.y = fe_neg(t.y);
            g1_add_mixed_ni(&acc, &acc, &t);
        }
        int b = d2[i];
        if (b) {
            int m = b < 0 ? -b : b;
            int idx = mode == 0 ? m - 1 : m >> 1;
            t.x = bx[idx]; t.y = TAB_Y(idx);
            if (b > 0) t.y = fe_neg(t.y);
            g1_add_mixed_ni(&acc, &acc, &t);
        }
    }
    acc.z = fp_mul(acc.z, zg);
    *out = acc;
#undef TAB_X
#undef TAB_Y
}

__host__ __device__ __forceinline__ void g1_mul_program(G1J* out, const G1J* p, const ScalarProgram* prog, Fp* ext = nullptr, unsigned estride = 1) {
    int top = prog->top;
    if (prog->is_one) { *out = *p; return; }
    g1_mul_digits(out, p, prog->d1, prog->d2, top, prog->mode, ext, estride);
}

// ---- on-device recoding of a variable scalar (canonical 8 x u32, < r) ---------------------
// k = k1 + k2 z^2 by long division with the 128-bit constant z^2, then fixed signed 4-bit
// windows (mode 0) for both halves.
struct LaneDigits {
    int8_t d1[B200_WNAF_LEN];
    int8_t d2[B200_WNAF_LEN];
    int top;
};

__host__ __device__ __forceinline__ bool u128_ge(const uint32_t a[4], const uint32_t b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > b[i]) return true;
        if (a[i] < b[i]) return false;
    }
    return true;
}
__host__ __device__ __forceinline__ void window4_recode(const uint32_t v_in[4], int8_t* out, int& top) {
    uint32_t v[5] = {v_in[0], v_in[1], v_in[2], v_in[3], 0};
    uint32_t carry = 0;
    for (int w = 0; w < 33; w++) {
        uint32_t d = ((v[w >> 3] >> ((w & 7) * 4)) & 15u) + carry;
        int dd;
        if (d > 8) { dd = (int)d - 16; carry = 1; } else { dd = (int)d; carry = 0; }
        out[4 * w] = (int8_t)dd;
        if (dd != 0 && 4 * w > top) top = 4 * w;
    }
}
static __host__ __device__ __noinline__ void g1_recode_scalar(LaneDigits* ld, const uint32_t* k) {
    constexpr uint32_t z2[4] = B200_GLV_Z2;
    uint32_t rem[4] = {0, 0, 0, 0}, quo[4] = {0, 0, 0, 0};
    for (int bit = 255; bit >= 0; bit--) {
        uint32_t topbit = rem[3] >> 31;
        rem[3] = (rem[3] << 1) | (rem[2] >> 31);
        rem[2] = (rem[2] << 1) | (rem[1] >> 31);
        rem[1] = (rem[1] << 1) | (rem[0] >> 31);
        rem[0] = (rem[0] << 1) | ((k[bit >> 5] >> (bit & 31)) & 1u);
        quo[3] = (quo[3] << 1) | (quo[2] >> 31);
        quo[2] = (quo[2] << 1) | (quo[1] >> 31);
        quo[1] = (quo[1] << 1) | (quo[0] >> 31);
        quo[0] = quo[0] << 1;
        if (topbit || u128_ge(rem, z2)) {
            uint32_t cf = 0;
            rem[0] = sub_cc(rem[0], z2[0], cf);
            rem[1] = subc_cc(rem[1], z2[1], cf);
            rem[2] = subc_cc(rem[2], z2[2], cf);
            rem[3] = subc(rem[3], z2[3], cf);
            quo[0] |= 1u;
        }
    }
    for (int i = 0; i < B200_WNAF_LEN; i++) { ld->d1[i] = 0; ld->d2[i] = 0; }
    int top = -1;
    window4_recode(rem, ld->d1, top);
    window4_recode(quo, ld->d2, top);
    ld->top = top;
}

// k * P for a per-lane scalar (canonical limbs)
__host__ __device__ __forceinline__ void g1_mul_var(G1J* out, const G1J* p, const uint32_t* k_canon) {
    LaneDigits ld;
    g1_recode_scalar(&ld, k_canon);
    g1_mul_digits(out, p, ld.d1, ld.d2, ld.top, 0);
}

}  // namespace b200
