// field_hybrid.cuh (included by field.cuh after field_fp64_impl.cuh and field_karatsuba.cuh)
//
// Montgomery product for Fp with the two halves on two different pipes INSIDE one call:
//
//     product half   a * b as 16 x 16 limbs of 24 bits in doubles: 256 exact DFMAs (136 for a square),
//                    one accumulation chain per column, no splitting, no m * p terms  -> FP64 pipe
//     glue           31 column sums (< 2^52, read straight out of the mantissa) -> 24 words -> ALU pipe
//     reduction half kara_redc: 12 rows of m * p only, 12 + 144 IMAD / IMAD.WIDE    -> FMA-heavy pipe
//
// field_fp64_impl.cuh ran the WHOLE product on the FP64 pipe (764 FP64 instructions) for some of the
// resident CTAs and lost; here every warp alternates between the pipes, so with 4 warps
// per sub-partition the three pipes overlap statistically.  Same R = 2^384, same canonical result as
// fe_mul, bit for bit (tools/mul_probe.cu compares them on the device, tests/test_abi_host.py on the host).
#pragma once

namespace b200 {

// one column of the 24-bit limb product; starts at 2^52 so that the sum (< 2^52) is the mantissa
template <bool SQR, int C>
HD double hyb_column(const double* a, const double* b) {
    constexpr int LO = C < 16 ? 0 : C - 15, HI = C < 16 ? C : 15;
    double s = 0x1p52;
    if (SQR) {
#pragma unroll
        for (int i = LO; i <= HI; i++) {
            if (i < C - i) s = fp64_fma(b[i], a[C - i], s);          // b = 2a: the doubled cross terms
            else if (i == C - i) s = fp64_fma(a[i], a[i], s);
        }
    } else {
#pragma unroll
        for (int i = LO; i <= HI; i++) s = fp64_fma(a[i], b[C - i], s);
    }
    return s;
}

HD void hyb_split(double s, uint32_t& lo, uint32_t& hi) {
#ifdef __CUDA_ARCH__
    lo = (uint32_t)__double2loint(s);
    hi = (uint32_t)__double2hiint(s) & 0xfffffu;
#else
    uint64_t v; memcpy(&v, &s, 8);
    lo = (uint32_t)v; hi = (uint32_t)(v >> 32) & 0xfffffu;
#endif
}

// Columns 4g .. 4g+3 (weights 2^0, 2^24, 2^48, 2^72 inside the group) summed into words 3g .. 3g+3 of
// T; word 3g already holds the previous group's top word.  G < 2^125, so nothing leaves word 3g+3.
template <bool SQR, int G>
struct HybGroups {
    static HD void run(uint32_t* T, const double* a, const double* b) {
        constexpr int C0 = 4 * G;
        uint32_t l0, h0, l1, h1, l2, h2, l3 = 0, h3 = 0;
        hyb_split(hyb_column<SQR, C0>(a, b), l0, h0);
        hyb_split(hyb_column<SQR, C0 + 1>(a, b), l1, h1);
        hyb_split(hyb_column<SQR, C0 + 2>(a, b), l2, h2);
        if (C0 + 3 < 31) hyb_split(hyb_column<SQR, C0 + 3>(a, b), l3, h3);
        uint32_t cf = 0, w0, w1, w2, w3;
        // v0 + carry word of the previous group
        w0 = add_cc(l0, G ? T[3 * G] : 0u, cf);
        w1 = addc(h0, 0u, cf);
        // + v1 << 24
        w0 = add_cc(w0, l1 << 24, cf);
        w1 = addc_cc(w1, (l1 >> 8) | (h1 << 24), cf);
        w2 = addc(h1 >> 8, 0u, cf);
        // + v2 << 48
        w1 = add_cc(w1, l2 << 16, cf);
        w2 = addc_cc(w2, (l2 >> 16) | (h2 << 16), cf);
        w3 = addc(h2 >> 16, 0u, cf);
        // + v3 << 72
        w2 = add_cc(w2, l3 << 8, cf);
        w3 = addc(w3, (l3 >> 24) | (h3 << 8), cf);
        T[3 * G] = w0; T[3 * G + 1] = w1; T[3 * G + 2] = w2;
        if (G < 7) T[3 * G + 3] = w3;
        HybGroups<SQR, G + 1>::run(T, a, b);
    }
};
template <bool SQR>
struct HybGroups<SQR, 8> {
    static HD void run(uint32_t*, const double*, const double*) {}
};

template <bool SQR>
HD Fp fe_mulsqr_hyb(const Fp& A, const Fp& B) {
    double a[16], b[16];
    fp64_expand(a, A.l);
    if (SQR) {
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = a[i] + a[i];
    } else {
        fp64_expand(b, B.l);
    }
    uint32_t T[24];
    HybGroups<SQR, 0>::run(T, a, b);
    return kara_redc<FpParams>(T);
}
// ---------------------------------------------------------------------------------------------------------------------
// Interleaved form.  The rows of the reduction are made independent of the high half of T and of all but ONE word of the
// low half: with U the running sum of the m_i p rows alone (T is NOT folded in),
//     m_r = (U_0 + T_r + c_r) * (-1/p)   mod 2^32,      U <- (U + m_r p) >> 32,      c_(r+1) = (T_r | c_r) != 0
// (the discarded low word of U + m_r p is -(T_r + c_r) mod 2^32; together with T_r + c_r it carries exactly when that
// is non-zero), and at the end   result = T_hi + U + c_12.   Row r therefore only waits for word r of the product, so the
// rows (FMA-heavy pipe) can be issued between the columns that produce the later words (FP64 pipe):
//     G0 | G1 + rows 0..2 | G2 + rows 3..5 | G3 + rows 6..8 | G4..G7 + rows 9..11 | T_hi + U
// Row arithmetic: the two-array even / odd carry chains of field_karatsuba.cuh (A word aligned, B the previous row's A).
template <int ROW>
HD void hyb_row(uint32_t* A, uint32_t* B, uint32_t Tr, uint32_t& cin) {
    typedef FpParams P;
    constexpr int N = P::N;
    uint32_t cf = 0;
    const uint32_t t0 = add_cc(A[0], B[1], cf);
    const uint32_t m = (t0 + Tr + cin) * P::INV;
#pragma unroll
    for (int j = 1; j < N; j += 2) {
        B[j - 1] = madc_lo_cc(m, P::mod(j), B[j + 1], cf);
        B[j] = madc_hi_cc(m, P::mod(j), (j + 2 <= N) ? B[j + 2] : 0u, cf);
    }
    B[N] = addc(0u, 0u, cf);
    A[0] = mad_lo_cc(m, P::mod(0), t0, cf);          // low word: -(Tr + cin), dropped; its carry is U's own
    A[1] = madc_hi_cc(m, P::mod(0), A[1], cf);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        A[j] = madc_lo_cc(m, P::mod(j), A[j], cf);
        A[j + 1] = madc_hi_cc(m, P::mod(j), A[j + 1], cf);
    }
    A[N] = addc(A[N], 0u, cf);
    cin = (Tr | cin) ? 1u : 0u;
}
// one group of four columns -> three final words of T (+ the carry word), as HybGroups but a single step
template <bool SQR, int G>
HD void hyb_group(uint32_t* T, const double* a, const double* b) {
    constexpr int C0 = 4 * G;
    uint32_t l0, h0, l1, h1, l2, h2, l3 = 0, h3 = 0;
    hyb_split(hyb_column<SQR, C0>(a, b), l0, h0);
    hyb_split(hyb_column<SQR, C0 + 1>(a, b), l1, h1);
    hyb_split(hyb_column<SQR, C0 + 2>(a, b), l2, h2);
    if (C0 + 3 < 31) hyb_split(hyb_column<SQR, C0 + 3>(a, b), l3, h3);
    uint32_t cf = 0, w0, w1, w2, w3;
    w0 = add_cc(l0, G ? T[3 * G] : 0u, cf);
    w1 = addc(h0, 0u, cf);
    w0 = add_cc(w0, l1 << 24, cf);
    w1 = addc_cc(w1, (l1 >> 8) | (h1 << 24), cf);
    w2 = addc(h1 >> 8, 0u, cf);
    w1 = add_cc(w1, l2 << 16, cf);
    w2 = addc_cc(w2, (l2 >> 16) | (h2 << 16), cf);
    w3 = addc(h2 >> 16, 0u, cf);
    w2 = add_cc(w2, l3 << 8, cf);
    w3 = addc(w3, (l3 >> 24) | (h3 << 8), cf);
    T[3 * G] = w0; T[3 * G + 1] = w1; T[3 * G + 2] = w2;
    if (G < 7) T[3 * G + 3] = w3;
}
template <bool SQR>
HD Fp fe_mulsqr_hyb2(const Fp& Ain, const Fp& Bin) {
    constexpr int N = 12;
    double a[16], b[16];
    fp64_expand(a, Ain.l);
    if (SQR) {
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = a[i] + a[i];
    } else {
        fp64_expand(b, Bin.l);
    }
    uint32_t T[24], X[N + 2], Y[N + 2], cin = 0;
#pragma unroll
    for (int k = 0; k < N + 2; k++) { X[k] = 0; Y[k] = 0; }
    hyb_group<SQR, 0>(T, a, b);
    hyb_group<SQR, 1>(T, a, b);
    hyb_row<0>(X, Y, T[0], cin); hyb_row<1>(Y, X, T[1], cin); hyb_row<2>(X, Y, T[2], cin);
    hyb_group<SQR, 2>(T, a, b);
    hyb_row<3>(Y, X, T[3], cin); hyb_row<4>(X, Y, T[4], cin); hyb_row<5>(Y, X, T[5], cin);
    hyb_group<SQR, 3>(T, a, b);
    hyb_row<6>(X, Y, T[6], cin); hyb_row<7>(Y, X, T[7], cin); hyb_row<8>(X, Y, T[8], cin);
    hyb_group<SQR, 4>(T, a, b);
    hyb_row<9>(Y, X, T[9], cin);
    hyb_group<SQR, 5>(T, a, b);
    hyb_row<10>(X, Y, T[10], cin);
    hyb_group<SQR, 6>(T, a, b);
    hyb_row<11>(Y, X, T[11], cin);
    hyb_group<SQR, 7>(T, a, b);
    // twelve rows: the last one ran with (A, B) = (Y, X), Y word aligned: U word k = Y[k + 1] + X[k]; result = U + T_hi + cin
    Fp r; uint32_t cf = 0;
    r.l[0] = add_cc(Y[1], X[0], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(Y[k + 1], X[k], cf);
    r.l[N - 1] = addc(Y[N], X[N - 1], cf);
    r.l[0] = add_cc(r.l[0], cin, cf);
#pragma unroll
    for (int k = 1; k < N; k++) r.l[k] = addc_cc(r.l[k], 0u, cf);
    r.l[0] = add_cc(r.l[0], T[N], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(r.l[k], T[N + k], cf);
    r.l[N - 1] = addc(r.l[N - 1], T[2 * N - 1], cf);
    fe_reduce_once(r);
    return r;
}
// ---------------------------------------------------------------------------------------------------------------------
// Interleaved form with the ORDER pinned.  Left to itself ptxas hoists all 256 DFMAs in front of the rows (two long
// single-pipe phases again).  The dropped low word of every row, z_r = (U + m_r p)_0 + T_r + c_r, is zero by construction
// but not provably so; it becomes the low mantissa word of the 2^52 the column chains start from, which makes the columns
// scheduled after row r data-dependent on it at no cost (a register move).  Rows are spread evenly over the DFMA stream:
// row r after ~256 (r + 1/2) / 12 DFMAs, never before the group that completes word r.
HD double hyb_bias(uint32_t dep) {      // 2^52 + dep, dep == 0 at run time
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)dep);
#else
    return 0x1p52 + (double)dep;
#endif
}
template <bool SQR, int C>
HD void hyb_col(uint32_t* cl, uint32_t* ch, const double* a, const double* b, uint32_t dep) {
    constexpr int LO = C < 16 ? 0 : C - 15, HI = C < 16 ? C : 15;
    double s = hyb_bias(dep);
    if (SQR) {
#pragma unroll
        for (int i = LO; i <= HI; i++) {
            if (i < C - i) s = fp64_fma(b[i], a[C - i], s);
            else if (i == C - i) s = fp64_fma(a[i], a[i], s);
        }
    } else {
#pragma unroll
        for (int i = LO; i <= HI; i++) s = fp64_fma(a[i], b[C - i], s);
    }
    hyb_split(s, cl[C], ch[C]);
}
template <int G>
HD void hyb_combine(uint32_t* T, const uint32_t* cl, const uint32_t* ch) {
    constexpr int C0 = 4 * G;
    const uint32_t l0 = cl[C0], h0 = ch[C0], l1 = cl[C0 + 1], h1 = ch[C0 + 1], l2 = cl[C0 + 2], h2 = ch[C0 + 2];
    const uint32_t l3 = C0 + 3 < 31 ? cl[C0 + 3 < 31 ? C0 + 3 : 0] : 0u, h3 = C0 + 3 < 31 ? ch[C0 + 3 < 31 ? C0 + 3 : 0] : 0u;
    uint32_t cf = 0, w0, w1, w2, w3;
    w0 = add_cc(l0, G ? T[3 * G] : 0u, cf);
    w1 = addc(h0, 0u, cf);
    w0 = add_cc(w0, l1 << 24, cf);
    w1 = addc_cc(w1, (l1 >> 8) | (h1 << 24), cf);
    w2 = addc(h1 >> 8, 0u, cf);
    w1 = add_cc(w1, l2 << 16, cf);
    w2 = addc_cc(w2, (l2 >> 16) | (h2 << 16), cf);
    w3 = addc(h2 >> 16, 0u, cf);
    w2 = add_cc(w2, l3 << 8, cf);
    w3 = addc(w3, (l3 >> 24) | (h3 << 8), cf);
    T[3 * G] = w0; T[3 * G + 1] = w1; T[3 * G + 2] = w2;
    if (G < 7) T[3 * G + 3] = w3;
}
// hyb_row that also hands out its (zero) low word
template <int ROW>
HD uint32_t hyb_row_z(uint32_t* A, uint32_t* B, uint32_t Tr, uint32_t& cin) {
    const uint32_t c0 = cin;
    hyb_row<ROW>(A, B, Tr, cin);
    return A[0] + Tr + c0;
}
template <bool SQR>
HD Fp fe_mulsqr_hyb3(const Fp& Ain, const Fp& Bin) {
    constexpr int N = 12;
    double a[16], b[16];
    fp64_expand(a, Ain.l);
    if (SQR) {
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = a[i] + a[i];
    } else {
        fp64_expand(b, Bin.l);
    }
    uint32_t T[24], X[N + 2], Y[N + 2], cl[31], ch[31], cin = 0, z = 0;
#pragma unroll
    for (int k = 0; k < N + 2; k++) { X[k] = 0; Y[k] = 0; }
#define COL(C) hyb_col<SQR, C>(cl, ch, a, b, z)
    COL(0); COL(1); COL(2); COL(3); hyb_combine<0>(T, cl, ch);
    z = hyb_row_z<0>(X, Y, T[0], cin);
    COL(4); COL(5); COL(6); COL(7); hyb_combine<1>(T, cl, ch);
    z = hyb_row_z<1>(Y, X, T[1], cin);
    COL(8); COL(9);
    z = hyb_row_z<2>(X, Y, T[2], cin);
    COL(10); COL(11); hyb_combine<2>(T, cl, ch);
    z = hyb_row_z<3>(Y, X, T[3], cin);
    COL(12); COL(13);
    z = hyb_row_z<4>(X, Y, T[4], cin);
    COL(14);
    z = hyb_row_z<5>(Y, X, T[5], cin);
    COL(15); hyb_combine<3>(T, cl, ch); COL(16);
    z = hyb_row_z<6>(X, Y, T[6], cin);
    COL(17);
    z = hyb_row_z<7>(Y, X, T[7], cin);
    COL(18); COL(19); hyb_combine<4>(T, cl, ch);
    z = hyb_row_z<8>(X, Y, T[8], cin);
    COL(20); COL(21);
    z = hyb_row_z<9>(Y, X, T[9], cin);
    COL(22); COL(23); hyb_combine<5>(T, cl, ch);
    z = hyb_row_z<10>(X, Y, T[10], cin);
    COL(24); COL(25); COL(26);
    z = hyb_row_z<11>(Y, X, T[11], cin);
    COL(27); hyb_combine<6>(T, cl, ch); COL(28); COL(29); COL(30); hyb_combine<7>(T, cl, ch);
#undef COL
    Fp r; uint32_t cf = 0;
    r.l[0] = add_cc(Y[1], X[0], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(Y[k + 1], X[k], cf);
    r.l[N - 1] = addc(Y[N], X[N - 1], cf);
    r.l[0] = add_cc(r.l[0], cin, cf);
#pragma unroll
    for (int k = 1; k < N; k++) r.l[k] = addc_cc(r.l[k], 0u, cf);
    r.l[0] = add_cc(r.l[0], T[N], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(r.l[k], T[N + k], cf);
    r.l[N - 1] = addc(r.l[N - 1], T[2 * N - 1], cf);
    fe_reduce_once(r);
    return r;
}
HD Fp fe_mul_hyb3(const Fp& a, const Fp& b) { return fe_mulsqr_hyb3<false>(a, b); }
HD Fp fe_sqr_hyb3(const Fp& a) { return fe_mulsqr_hyb3<true>(a, a); }

HD Fp fe_mul_hyb2(const Fp& a, const Fp& b) { return fe_mulsqr_hyb2<false>(a, b); }
HD Fp fe_sqr_hyb2(const Fp& a) { return fe_mulsqr_hyb2<true>(a, a); }

HD Fp fe_mul_hyb(const Fp& a, const Fp& b) { return fe_mulsqr_hyb<false>(a, b); }
HD Fp fe_sqr_hyb(const Fp& a) { return fe_mulsqr_hyb<true>(a, a); }

}  // namespace b200
