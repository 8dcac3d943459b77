// field_karatsuba.cuh (included by field.cuh once fe_reduce_once and the carry primitives exist)
//
// Montgomery product / square for N = 12 limbs with the product half done by Karatsuba and the
// reduction half by word-serial REDC rows that only carry the m * p products:
//
//     fe_mul  (CIOS, field.cuh)      144 + 144 + 12 = 300 multiply-adds
//     fe_mul_k  (here)               3 * 36 (one level; 2 * 27 + ... with KARA_LEVELS 2) + 144 + 12
//     fe_sqr  (field.cuh)             78 + 144 + 12 = 234
//     fe_sqr_k  (here)               3 * 21 + 144 + 12 = 219
//
// The G1 kernels are bound by the issue rate of IMAD.WIDE (FMA-heavy pipe 84 % busy, ALU pipe 24 %,
// profiles/r01_stage_kernel_batch128.txt), so trading multiply-adds for additions on the ALU pipe pays.
// Same carry-chain style as field.cuh (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.X;
// the host build emulates the carry flag), same canonical result.
#pragma once

namespace b200 {

// r[0..2N) = a[0..N) * b[0..N): schoolbook.  Products whose low word lands on an even word go to E
// (E[k] is word k), the others to O (O[k] is word k + 1), so that every row is two carry chains over
// disjoint (lo, hi) register pairs.
template <int N>
HD void kara_mul_sb(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t E[2 * N], O[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t cf = 0;
        {
            bool first = true; int top = 0;
#pragma unroll
            for (int j = (i & 1); j < N; j += 2) {
                const int p = i + j;
                E[p] = first ? mad_lo_cc(a[j], b[i], E[p], cf) : madc_lo_cc(a[j], b[i], E[p], cf);
                E[p + 1] = madc_hi_cc(a[j], b[i], E[p + 1], cf);
                first = false; top = p + 2;
            }
            if (!first && top < 2 * N) E[top] = addc(E[top], 0u, cf);
        }
        {
            bool first = true; int top = 0;
#pragma unroll
            for (int j = 1 - (i & 1); j < N; j += 2) {
                const int p = i + j - 1;
                O[p] = first ? mad_lo_cc(a[j], b[i], O[p], cf) : madc_lo_cc(a[j], b[i], O[p], cf);
                O[p + 1] = madc_hi_cc(a[j], b[i], O[p + 1], cf);
                first = false; top = p + 2;
            }
            if (!first && top < 2 * N) O[top] = addc(O[top], 0u, cf);
        }
    }
    uint32_t cf = 0;
    r[0] = E[0];
    r[1] = add_cc(E[1], O[0], cf);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) r[k] = addc_cc(E[k], O[k - 1], cf);
    r[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2], cf);
}

// r[0..2N) = a[0..N)^2: cross products once (E / O chains as above), doubled, plus the diagonal.
template <int N>
HD void kara_sqr_sb(uint32_t* r, const uint32_t* a) {
    uint32_t E[2 * N], O[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < N - 1; i++) {              // row i: a_j * a_i for j > i
        uint32_t cf = 0;
        {
            bool first = true; int top = 0;
#pragma unroll
            for (int j = i + 2; j < N; j += 2) {    // i + j even
                const int p = i + j;
                E[p] = first ? mad_lo_cc(a[j], a[i], E[p], cf) : madc_lo_cc(a[j], a[i], E[p], cf);
                E[p + 1] = madc_hi_cc(a[j], a[i], E[p + 1], cf);
                first = false; top = p + 2;
            }
            if (!first && top < 2 * N) E[top] = addc(E[top], 0u, cf);
        }
        {
            bool first = true; int top = 0;
#pragma unroll
            for (int j = i + 1; j < N; j += 2) {    // i + j odd
                const int p = i + j - 1;
                O[p] = first ? mad_lo_cc(a[j], a[i], O[p], cf) : madc_lo_cc(a[j], a[i], O[p], cf);
                O[p + 1] = madc_hi_cc(a[j], a[i], O[p + 1], cf);
                first = false; top = p + 2;
            }
            if (!first && top < 2 * N) O[top] = addc(O[top], 0u, cf);
        }
    }
    // cross = E + W O; r = 2 cross
    uint32_t c[2 * N], cf = 0;
    c[0] = E[0];
    c[1] = add_cc(E[1], O[0], cf);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; k++) c[k] = addc_cc(E[k], O[k - 1], cf);
    c[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2], cf);
    r[0] = c[0] << 1;
#pragma unroll
    for (int k = 1; k < 2 * N; k++) r[k] = (c[k] << 1) | (c[k - 1] >> 31);
    // + diagonal a_i^2 at word 2 i: one chain over all 2N words
    r[0] = mad_lo_cc(a[0], a[0], r[0], cf);
    r[1] = madc_hi_cc(a[0], a[0], r[1], cf);
#pragma unroll
    for (int i = 1; i < N; i++) {
        r[2 * i] = madc_lo_cc(a[i], a[i], r[2 * i], cf);
        r[2 * i + 1] = (i == N - 1) ? madc_hi(a[i], a[i], r[2 * i + 1], cf) : madc_hi_cc(a[i], a[i], r[2 * i + 1], cf);
    }
}

// d = |x - y| over H words; returns the sign mask (all ones if x < y)
template <int H>
HD uint32_t kara_abs_diff(uint32_t* d, const uint32_t* x, const uint32_t* y) {
    uint32_t cf = 0;
    d[0] = sub_cc(x[0], y[0], cf);
#pragma unroll
    for (int k = 1; k < H; k++) d[k] = subc_cc(x[k], y[k], cf);
    const uint32_t mask = subc(0u, 0u, cf);
    d[0] = add_cc(d[0] ^ mask, mask & 1u, cf);
#pragma unroll
    for (int k = 1; k < H; k++) d[k] = (k == H - 1) ? addc(d[k] ^ mask, 0u, cf) : addc_cc(d[k] ^ mask, 0u, cf);
    return mask;
}

// r[h .. ) += mid (2h + 1 words: z0 + z2 +- zm), z0 = r[0..2h), z2 = r[2h..4h)
template <int H>
HD void kara_fold_mid(uint32_t* r, const uint32_t* zm, uint32_t neg_mask) {
    uint32_t mid[2 * H + 1], cf = 0;
    mid[0] = add_cc(r[0], r[2 * H], cf);
#pragma unroll
    for (int k = 1; k < 2 * H; k++) mid[k] = addc_cc(r[k], r[2 * H + k], cf);
    mid[2 * H] = addc(0u, 0u, cf);
    // mid += neg ? -zm : zm   (two's complement over 2h + 1 words; the result is non-negative)
    (void)add_cc(neg_mask, neg_mask, cf);                 // carry = neg_mask & 1
#pragma unroll
    for (int k = 0; k < 2 * H; k++) mid[k] = addc_cc(mid[k], zm[k] ^ neg_mask, cf);
    mid[2 * H] = addc(mid[2 * H], neg_mask, cf);
    r[H] = add_cc(r[H], mid[0], cf);
#pragma unroll
    for (int k = 1; k <= 2 * H; k++) r[H + k] = addc_cc(r[H + k], mid[k], cf);
#pragma unroll
    for (int k = 3 * H + 1; k < 4 * H; k++) r[k] = (k == 4 * H - 1) ? addc(r[k], 0u, cf) : addc_cc(r[k], 0u, cf);
}

#ifndef B200_KARA_LEVELS
#define B200_KARA_LEVELS 1
#endif
// r[0..2N) = a * b, N = 12 (or 6 inside the second level)
template <int N, int LEVELS>
struct KaraMul {
    static HD void run(uint32_t* r, const uint32_t* a, const uint32_t* b) {
        constexpr int H = N / 2;
        uint32_t da[H], db[H], zm[2 * H];
        const uint32_t sa = kara_abs_diff<H>(da, a, a + H);             // a_lo - a_hi
        const uint32_t sb = kara_abs_diff<H>(db, b + H, b);             // b_hi - b_lo
        KaraMul<H, LEVELS - 1>::run(r, a, b);                           // z0
        KaraMul<H, LEVELS - 1>::run(r + 2 * H, a + H, b + H);           // z2
        KaraMul<H, LEVELS - 1>::run(zm, da, db);                        // |a_lo - a_hi| |b_hi - b_lo|
        kara_fold_mid<H>(r, zm, sa ^ sb);                               // a_lo b_hi + a_hi b_lo = z0 + z2 + (a_lo - a_hi)(b_hi - b_lo)
    }
};
template <int N>
struct KaraMul<N, 0> {
    static HD void run(uint32_t* r, const uint32_t* a, const uint32_t* b) { kara_mul_sb<N>(r, a, b); }
};
template <int N, int LEVELS>
HD void kara_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) { KaraMul<N, LEVELS>::run(r, a, b); }
// r[0..2N) = a^2: 2 a_lo a_hi = a_lo^2 + a_hi^2 - (a_lo - a_hi)^2
template <int N>
HD void kara_sqr(uint32_t* r, const uint32_t* a) {
    constexpr int H = N / 2;
    uint32_t d[H], zm[2 * H];
    (void)kara_abs_diff<H>(d, a, a + H);
    kara_sqr_sb<H>(r, a);
    kara_sqr_sb<H>(r + 2 * H, a + H);
    kara_sqr_sb<H>(zm, d);
    kara_fold_mid<H>(r, zm, 0xffffffffu);
}

// Montgomery reduction of T (2N words, T < p W^N): N rows of m * p only.  Same two-array scheme as
// mont_row in field.cuh -- A is word aligned, B is the previous row's A, which moves down two words
// while the odd products are added, its word 1 being the pending low word -- with the high words of T
// fed in at the top of the odd chain, one per row.
template <class P, int ROW>
struct KaraRedcRows {
    static HD void run(uint32_t* A, uint32_t* B, const uint32_t* T) {
        constexpr int N = P::N;
        uint32_t cf = 0;
        const uint32_t t0 = add_cc(A[0], B[1], cf);
        const uint32_t m = t0 * P::INV;
#pragma unroll
        for (int j = 1; j < N; j += 2) {
            B[j - 1] = madc_lo_cc(m, P::mod(j), B[j + 1], cf);
            B[j] = madc_hi_cc(m, P::mod(j), (j + 2 <= N) ? B[j + 2] : T[N + ROW], cf);
        }
        B[N] = addc(0u, 0u, cf);
        A[0] = mad_lo_cc(m, P::mod(0), t0, cf);
        A[1] = madc_hi_cc(m, P::mod(0), A[1], cf);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            A[j] = madc_lo_cc(m, P::mod(j), A[j], cf);
            A[j + 1] = madc_hi_cc(m, P::mod(j), A[j + 1], cf);
        }
        A[N] = addc(A[N], 0u, cf);
        KaraRedcRows<P, ROW + 1>::run(B, A, T);
    }
};
template <class P>
struct KaraRedcRows<P, P::N> {
    static HD void run(uint32_t*, uint32_t*, const uint32_t*) {}
};
template <class P>
HD Fe<P> kara_redc(const uint32_t* T) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t A[N + 1], B[N + 2];
#pragma unroll
    for (int k = 0; k < N; k++) { A[k] = T[k]; B[k] = 0; }
    A[N] = 0; B[N] = 0; B[N + 1] = 0;
    KaraRedcRows<P, 0>::run(A, B, T);
    // N rows (even count): the last row ran with the roles swapped -- B word aligned with its word 0
    // cleared, A holding the odd chain.  After the last division by W: result word k = B[k + 1] + A[k].
    Fe<P> r; uint32_t cf = 0;
    r.l[0] = add_cc(B[1], A[0], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(B[k + 1], A[k], cf);
    r.l[N - 1] = addc(B[N], A[N - 1], cf);
    fe_reduce_once(r);
    return r;
}

template <class P>
HD Fe<P> fe_mul_k(const Fe<P>& a, const Fe<P>& b) {
    uint32_t T[2 * P::N];
    kara_mul<P::N, B200_KARA_LEVELS>(T, a.l, b.l);
    return kara_redc<P>(T);
}
template <class P>
HD Fe<P> fe_sqr_k(const Fe<P>& a) {
    uint32_t T[2 * P::N];
    kara_sqr<P::N>(T, a.l);
    return kara_redc<P>(T);
}

}  // namespace b200
