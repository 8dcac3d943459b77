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
HD Fp fe_mul_hyb(const Fp& a, const Fp& b) { return fe_mulsqr_hyb<false>(a, b); }
HD Fp fe_sqr_hyb(const Fp& a) { return fe_mulsqr_hyb<true>(a, a); }

}  // namespace b200
