// field_fp64_impl.cuh (included by field.cuh once Fp is defined) -- the 381-bit Montgomery product of field.cuh computed on the FP64 pipe.
//
// Why: the G1 kernels are bound by the issue rate of IMAD.WIDE (one per 4 cycles per SM
// sub-partition, profiles/r01_pipe_probe.txt) while the FP64 pipe of a B200 (one DFMA per
// 2 cycles per sub-partition, issued concurrently with IMAD.WIDE) idles.  Warps that run this
// routine instead of fe_mul add multiplier throughput instead of competing for it.
//
// How: operands are split into 16 limbs of 24 bits held in doubles.  A limb product is < 2^48, so
// a column of the schoolbook product -- at most 16 a_i b_j terms, 16 m_i p_j terms and one carry
// -- stays below 2^53: every DFMA below is EXACT, no rounding ever happens.  The only rounding
// dependent step is floor(x / 2^24), done with the round-towards-zero FMA against 2^52.  Column by
// column (product scanning), with the Montgomery quotient digit m_c = -t p^-1 mod 2^24 taken as
// soon as column c < 16 is complete: R = 2^(24*16) = 2^384, the same R as the 12 x 32-bit integer
// routine, hence the same (canonical) result bit for bit.
//
// __host__ __device__ like the rest of the field code: the host build replaces the two intrinsics
// by floor() / fma(), so the exact routine is unit-tested on the CPU against fe_mul.
#pragma once
#include <math.h>

namespace b200 {

#ifdef __CUDACC__
// p's limbs sit in constant memory on the device: with a compile-time index they become direct
// c[bank][offset] operands of the DFMA (as immediates every use would cost two MOVs).
static __constant__ double k_fp64_p24[16] = B200_FP_MOD24;
#endif
HD double fp64_p24(int i) {
#ifdef __CUDA_ARCH__
    return k_fp64_p24[i];
#else
    constexpr double t[16] = B200_FP_MOD24; return t[i];
#endif
}

HD double fp64_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
// floor(x / 2^24) for an integer-valued 0 <= x < 2^53
HD double fp64_floor24(double x) {
#ifdef __CUDA_ARCH__
    return __fma_rz(x, 0x1p-24, 0x1p52) - 0x1p52;
#else
    return floor(x * 0x1p-24);
#endif
}
HD double fp64_from_u32(uint32_t v) {   // v < 2^32, exact
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)v) - 0x1p52;
#else
    return (double)v;
#endif
}
HD uint32_t fp64_to_u32(double x) {     // integer-valued 0 <= x < 2^32, exact
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(x + 0x1p52);
#else
    return (uint32_t)x;
#endif
}

// 12 x 32-bit words -> 16 x 24-bit limbs (as doubles)
HD void fp64_expand(double* d, const uint32_t* w) {
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
        d[4 * g + 0] = fp64_from_u32(w0 & 0xffffffu);
        d[4 * g + 1] = fp64_from_u32(((w0 >> 24) | (w1 << 8)) & 0xffffffu);
        d[4 * g + 2] = fp64_from_u32(((w1 >> 16) | (w2 << 16)) & 0xffffffu);
        d[4 * g + 3] = fp64_from_u32(w2 >> 8);
    }
}

// One column of the product scanning loop, C a compile-time constant so that every inner trip
// count and array index is static (ptxas keeps a, b, m in registers; left to `#pragma unroll` on a
// run-time column index nvcc kept the late columns as loops over local-memory arrays).
template <bool SQR, int C>
HD void fp64_column(const double* a, const double* b, double* m, double* r, double& carry) {
    constexpr int LO = C < 16 ? 0 : C - 15, HI = C < 16 ? C : 15;
    double s = carry, u = 0.0;           // two independent accumulation chains: a*b and m*p
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
#pragma unroll
    for (int i = LO; i <= HI; i++)
        if (i != C) u = fp64_fma(m[i], fp64_p24(C - i), u);
    double t = s + u;
    if (C < 16) {
        double q = fp64_floor24(t);
        double l = fp64_fma(q, -0x1p24, t);                           // t mod 2^24
        double pr = l * (double)B200_FP_INV24;                        // < 2^48
        double q2 = fp64_floor24(pr);
        double mc = fp64_fma(q2, -0x1p24, pr);                        // m_C = -t / p mod 2^24
        m[C] = mc;
        t = fp64_fma(mc, fp64_p24(0), t);                             // now divisible by 2^24
        carry = t * 0x1p-24;
    } else {
        double q = fp64_floor24(t);
        r[C - 16] = fp64_fma(q, -0x1p24, t);
        carry = q;
    }
}
template <bool SQR, int C>
struct Fp64Columns {
    static HD void run(const double* a, const double* b, double* m, double* r, double& carry) {
        fp64_column<SQR, C>(a, b, m, r, carry);
        Fp64Columns<SQR, C + 1>::run(a, b, m, r, carry);
    }
};
template <bool SQR>
struct Fp64Columns<SQR, 31> {
    static HD void run(const double*, const double*, double*, double*, double&) {}
};

// Montgomery product (SQR: square of a; b ignored).  a, b < p.
template <bool SQR>
HD Fp fe_mulsqr_fp64(const Fp& A, const Fp& B) {
    double a[16], b[16], m[16], r[16];
    fp64_expand(a, A.l);
    if (SQR) {
#pragma unroll
        for (int i = 0; i < 16; i++) b[i] = a[i] + a[i];      // doubled copy for the cross terms
    } else {
        fp64_expand(b, B.l);
    }
    double carry = 0.0;
    Fp64Columns<SQR, 0>::run(a, b, m, r, carry);
    r[15] = carry;
    Fp out;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t l0 = fp64_to_u32(r[4 * g]), l1 = fp64_to_u32(r[4 * g + 1]), l2 = fp64_to_u32(r[4 * g + 2]), l3 = fp64_to_u32(r[4 * g + 3]);
        out.l[3 * g + 0] = l0 | (l1 << 24);
        out.l[3 * g + 1] = (l1 >> 8) | (l2 << 16);
        out.l[3 * g + 2] = (l2 >> 16) | (l3 << 8);
    }
    fe_reduce_once(out);
    return out;
}
HD Fp fe_mul_fp64(const Fp& a, const Fp& b) { return fe_mulsqr_fp64<false>(a, b); }
HD Fp fe_sqr_fp64(const Fp& a) { return fe_mulsqr_fp64<true>(a, a); }

}  // namespace b200
