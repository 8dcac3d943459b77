// field.cuh -- Montgomery arithmetic over the two BLS12-381 primes on 32-bit limbs.
//
// One element per lane: Fr = 8 x u32 (R = 2^256), Fp = 12 x u32 (R = 2^384).
// The multiplication is an interleaved (CIOS-style) Montgomery product organised as two
// accumulators ("even"- and "odd"-aligned 64-bit product columns) so that every
// lo/hi product pair lands on a contiguous carry chain: mad.lo.cc / madc.hi.cc.  ptxas
// pairs those into IMAD.WIDE + carry predicates on sm_100a.
//
// Everything is __host__ __device__: on the device the carry primitives are inline PTX,
// on the host the same algorithm runs against an explicit emulated carry flag, so the
// exact code the kernels execute is unit-tested on the CPU (level-1 API, tests -m "not gpu").
//
// Replaces (semantics): kilic Fr/Fe arithmetic reached through bls/bignum_kilic.go:95-115
// and bls/bls_kilic.go:41-56 of the reference.
#pragma once
#include <stdint.h>
#include <string.h>
#include "constants.cuh"

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#define HDNI __host__ __device__ __noinline__
#else
#define HD inline
#define HDNI
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------
// carry-chain primitives.  `cf` is the emulated CC.CF on the host; unused on the device.
// ---------------------------------------------------------------------------------------
HD uint32_t add_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t s = (uint64_t)a + b; cf = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
HD uint32_t addc_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t s = (uint64_t)a + b + cf; cf = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
HD uint32_t addc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a + b + cf;
#endif
}
HD uint32_t sub_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t s = (uint64_t)a - b; cf = (uint32_t)(s >> 63); return (uint32_t)s;
#endif
}
HD uint32_t subc_cc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    uint64_t s = (uint64_t)a - b - cf; cf = (uint32_t)(s >> 63); return (uint32_t)s;
#endif
}
HD uint32_t subc(uint32_t a, uint32_t b, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
    return a - b - cf;
#endif
}
HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t s = (uint64_t)(a * b) + c; cf = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t s = (uint64_t)(a * b) + c + cf; cf = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    uint64_t s = (((uint64_t)a * b) >> 32) + c + cf; cf = (uint32_t)(s >> 32); return (uint32_t)s;
#endif
}
HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c, uint32_t& cf) {
#ifdef __CUDA_ARCH__
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32) + c + cf;
#endif
}

// ---------------------------------------------------------------------------------------
// field parameters
// ---------------------------------------------------------------------------------------
struct FpParams {
    static constexpr int N = B200_FP_LIMBS;
    static constexpr bool FAST_SQR = true;    // 3 p < 2^384: the doubled cross terms of fe_sqr fit the row accumulators
    static constexpr uint32_t INV = B200_FP_INV32;
    static HD constexpr uint32_t mod(int i) { constexpr uint32_t t[N] = B200_FP_MOD; return t[i]; }
    static HD constexpr uint32_t one(int i) { constexpr uint32_t t[N] = B200_FP_ONE; return t[i]; }
    static HD constexpr uint32_t r2(int i) { constexpr uint32_t t[N] = B200_FP_R2; return t[i]; }
    static HD constexpr uint32_t modm2(int i) { constexpr uint32_t t[N] = B200_FP_MODM2; return t[i]; }
};
struct FrParams {
    static constexpr int N = B200_FR_LIMBS;
    static constexpr bool FAST_SQR = false;   // 3 r > 2^256: squares go through fe_mul
    static constexpr uint32_t INV = B200_FR_INV32;
    static HD constexpr uint32_t mod(int i) { constexpr uint32_t t[N] = B200_FR_MOD; return t[i]; }
    static HD constexpr uint32_t one(int i) { constexpr uint32_t t[N] = B200_FR_ONE; return t[i]; }
    static HD constexpr uint32_t r2(int i) { constexpr uint32_t t[N] = B200_FR_R2; return t[i]; }
    static HD constexpr uint32_t modm2(int i) { constexpr uint32_t t[N] = B200_FR_MODM2; return t[i]; }
};

// ---------------------------------------------------------------------------------------
// Field element (value semantics; limbs little-endian). Montgomery or canonical content is
// the caller's business: add/sub are representation agnostic, mul divides by R.
// ---------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {   // 16 B alignment: elements move as uint4 vectors, also through local memory
    static constexpr int N = P::N;
    uint32_t l[N];

    static HD Fe zero() { Fe r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    static HD Fe one() { Fe r; for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }   // R mod p
    static HD Fe r2() { Fe r; for (int i = 0; i < N; i++) r.l[i] = P::r2(i); return r; }
    static HD Fe modulus() { Fe r; for (int i = 0; i < N; i++) r.l[i] = P::mod(i); return r; }

    HD bool is_zero() const { uint32_t a = 0; for (int i = 0; i < N; i++) a |= l[i]; return a == 0; }
    HD bool operator==(const Fe& o) const { uint32_t a = 0; for (int i = 0; i < N; i++) a |= l[i] ^ o.l[i]; return a == 0; }
    HD bool operator!=(const Fe& o) const { return !(*this == o); }
};

// r = (a >= p) ? a - p : a     (a < 2p)
template <class P>
HD void fe_reduce_once(Fe<P>& a) {
    constexpr int N = P::N;
    uint32_t t[N], cf = 0;
    t[0] = sub_cc(a.l[0], P::mod(0), cf);
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(a.l[i], P::mod(i), cf);
    uint32_t borrow = subc(0u, 0u, cf);   // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t[i];
}

template <class P>
HD Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r; uint32_t cf = 0;
    r.l[0] = add_cc(a.l[0], b.l[0], cf);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i], cf);
    r.l[N - 1] = addc(a.l[N - 1], b.l[N - 1], cf);   // both primes leave a spare top bit
    fe_reduce_once(r);
    return r;
}

template <class P>
HD Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    Fe<P> r; uint32_t cf = 0;
    r.l[0] = sub_cc(a.l[0], b.l[0], cf);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(a.l[i], b.l[i], cf);
    uint32_t borrow = subc(0u, 0u, cf);   // all ones if a < b
    uint32_t c2 = 0;
    r.l[0] = add_cc(r.l[0], P::mod(0) & borrow, c2);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(r.l[i], P::mod(i) & borrow, c2);
    r.l[N - 1] = addc(r.l[N - 1], P::mod(N - 1) & borrow, c2);
    return r;
}

template <class P>
HD Fe<P> fe_neg(const Fe<P>& a) {
    if (a.is_zero()) return a;
    return fe_sub(Fe<P>::zero(), a);
}
template <class P>
HD Fe<P> fe_dbl(const Fe<P>& a) { return fe_add(a, a); }

// ---------------------------------------------------------------------------------------
// Montgomery product, portable form (64-bit accumulation).  Cross-check for the PTX form.
// ---------------------------------------------------------------------------------------
template <class P>
HD Fe<P> fe_mul_portable(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    uint32_t t[N + 2];
    for (int i = 0; i < N + 2; i++) t[i] = 0;
    for (int i = 0; i < N; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < N; j++) {
            uint64_t s = (uint64_t)a.l[j] * b.l[i] + t[j] + carry;
            t[j] = (uint32_t)s; carry = s >> 32;
        }
        uint64_t s = (uint64_t)t[N] + carry;
        t[N] = (uint32_t)s; t[N + 1] = (uint32_t)(s >> 32);
        uint32_t m = t[0] * P::INV;
        s = (uint64_t)m * P::mod(0) + t[0];
        carry = s >> 32;
        for (int j = 1; j < N; j++) {
            s = (uint64_t)m * P::mod(j) + t[j] + carry;
            t[j - 1] = (uint32_t)s; carry = s >> 32;
        }
        s = (uint64_t)t[N] + carry;
        t[N - 1] = (uint32_t)s;
        t[N] = t[N + 1] + (uint32_t)(s >> 32);
    }
    Fe<P> r;
    for (int i = 0; i < N; i++) r.l[i] = t[i];
    fe_reduce_once(r);
    return r;
}

// ---------------------------------------------------------------------------------------
// Montgomery product, even/odd carry-chain form (the hot routine).
//
// T = X + W*Y (W = 2^32): X[k] sits at word k, Y[k] at word k+1.  One row adds a*b_i, then
// m*p with m = X[0]*INV, which clears word 0; the shift by one word is realised by swapping
// the roles of the two arrays for the next row (the old X, minus its two lowest words,
// becomes the new Y).  T < W^(N+1) throughout, so the carry out of an X chain is absorbed
// by Y[N-1] and the carry out of a Y chain is zero.
// ---------------------------------------------------------------------------------------
template <class P, bool FIRST>
HD void mont_row(uint32_t* X, uint32_t* Y, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    uint32_t cf = 0;
    if (FIRST) {
        // X = even products, Y = odd products (both arrays start at zero)
#pragma unroll
        for (int j = 0; j < N; j += 2) { X[j] = mul_lo(a[j], bi); X[j + 1] = mul_hi(a[j], bi); }
#pragma unroll
        for (int j = 1; j < N; j += 2) { Y[j - 1] = mul_lo(a[j], bi); Y[j] = mul_hi(a[j], bi); }
    } else {
        // here X is the previous row's Y (already word-aligned) and Y the previous row's X,
        // whose word 1 is still pending at word 0 and whose words 2.. become Y'[0..].
        X[0] = add_cc(X[0], Y[1], cf);
#pragma unroll
        for (int j = 1; j < N - 1; j += 2) {
            Y[j - 1] = madc_lo_cc(a[j], bi, Y[j + 1], cf);
            Y[j] = madc_hi_cc(a[j], bi, Y[j + 2], cf);
        }
        Y[N - 2] = madc_lo_cc(a[N - 1], bi, 0u, cf);
        Y[N - 1] = madc_hi(a[N - 1], bi, 0u, cf);
        X[0] = mad_lo_cc(a[0], bi, X[0], cf);
        X[1] = madc_hi_cc(a[0], bi, X[1], cf);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            X[j] = madc_lo_cc(a[j], bi, X[j], cf);
            X[j + 1] = madc_hi_cc(a[j], bi, X[j + 1], cf);
        }
        Y[N - 1] = addc(Y[N - 1], 0u, cf);
    }
    uint32_t m = X[0] * P::INV;
    X[0] = mad_lo_cc(m, P::mod(0), X[0], cf);
    X[1] = madc_hi_cc(m, P::mod(0), X[1], cf);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(m, P::mod(j), X[j], cf);
        X[j + 1] = madc_hi_cc(m, P::mod(j), X[j + 1], cf);
    }
    Y[N - 1] = addc(Y[N - 1], 0u, cf);
    Y[0] = mad_lo_cc(m, P::mod(1), Y[0], cf);
    Y[1] = madc_hi_cc(m, P::mod(1), Y[1], cf);
#pragma unroll
    for (int j = 3; j < N; j += 2) {
        Y[j - 1] = madc_lo_cc(m, P::mod(j), Y[j - 1], cf);
        Y[j] = madc_hi_cc(m, P::mod(j), Y[j], cf);
    }
}

template <class P>
HD Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t ev[N], od[N];
    mont_row<P, true>(ev, od, a.l, b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_row<P, false>(od, ev, a.l, b.l[i]);
        if (i + 1 < N) mont_row<P, false>(ev, od, a.l, b.l[i + 1]);
    }
    // last row had X = od (od[0] == 0 now): result word k = od[k+1] + ev[k]
    Fe<P> r; uint32_t cf = 0;
    r.l[0] = add_cc(od[1], ev[0], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(od[k + 1], ev[k], cf);
    r.l[N - 1] = addc(ev[N - 1], 0u, cf);
    fe_reduce_once(r);
    return r;
}

// ---------------------------------------------------------------------------------------
// Montgomery square: the same interleaved rows, but row J only multiplies the limbs j >= J of
//     c^(J) = { a_J, 2 a_(J+1) mod W, d_(J+2), .., d_(N-1) },   d_i = (a_i << 1) | (a_(i-1) >> 31)
// by a_J, i.e. a_J^2 plus the doubled cross terms a_J a_i (i > J) -- each cross product once
// (N (N + 1) / 2 products instead of N^2; the reduction half of a row is unchanged).  Skipped
// products degenerate into carry-propagating adds, which run on the otherwise idle ALU pipe.
// Row accumulators hold up to (2a + p) W, so this needs 3p < W^N: true for Fp (381 of 384 bits),
// not for Fr (255 of 256 bits), see FAST_SQR.
// ---------------------------------------------------------------------------------------
template <class P, int J>
HD uint32_t sq_operand(const uint32_t* a, const uint32_t* d, int j) {
    return j == J ? a[j] : (j == J + 1 ? (a[j] << 1) : d[j]);
}
template <class P, int J>
HD void mont_row_sq(uint32_t* X, uint32_t* Y, const uint32_t* a, const uint32_t* d) {
    constexpr int N = P::N;
    constexpr int FIRST_EVEN = (J + 1) & ~1;          // first even limb index >= J
    const uint32_t bi = a[J];
    uint32_t cf = 0;
    if (J == 0) {
#pragma unroll
        for (int j = 0; j < N; j += 2) { uint32_t c = sq_operand<P, J>(a, d, j); X[j] = mul_lo(c, bi); X[j + 1] = mul_hi(c, bi); }
#pragma unroll
        for (int j = 1; j < N; j += 2) { uint32_t c = sq_operand<P, J>(a, d, j); Y[j - 1] = mul_lo(c, bi); Y[j] = mul_hi(c, bi); }
    } else {
        X[0] = add_cc(X[0], Y[1], cf);
#pragma unroll
        for (int j = 1; j < N - 1; j += 2) {
            if (j >= J) {
                uint32_t c = sq_operand<P, J>(a, d, j);
                Y[j - 1] = madc_lo_cc(c, bi, Y[j + 1], cf);
                Y[j] = madc_hi_cc(c, bi, Y[j + 2], cf);
            } else {
                Y[j - 1] = addc_cc(Y[j + 1], 0u, cf);
                Y[j] = addc_cc(Y[j + 2], 0u, cf);
            }
        }
        {
            uint32_t c = sq_operand<P, J>(a, d, N - 1);
            Y[N - 2] = madc_lo_cc(c, bi, 0u, cf);
            Y[N - 1] = madc_hi(c, bi, 0u, cf);
        }
        if (FIRST_EVEN < N) {
#pragma unroll
            for (int j = 0; j < N; j += 2) {
                if (j < FIRST_EVEN) continue;
                uint32_t c = sq_operand<P, J>(a, d, j);
                if (j == FIRST_EVEN) { X[j] = mad_lo_cc(c, bi, X[j], cf); X[j + 1] = madc_hi_cc(c, bi, X[j + 1], cf); }
                else { X[j] = madc_lo_cc(c, bi, X[j], cf); X[j + 1] = madc_hi_cc(c, bi, X[j + 1], cf); }
            }
            Y[N - 1] = addc(Y[N - 1], 0u, cf);
        }
    }
    uint32_t m = X[0] * P::INV;
    X[0] = mad_lo_cc(m, P::mod(0), X[0], cf);
    X[1] = madc_hi_cc(m, P::mod(0), X[1], cf);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        X[j] = madc_lo_cc(m, P::mod(j), X[j], cf);
        X[j + 1] = madc_hi_cc(m, P::mod(j), X[j + 1], cf);
    }
    Y[N - 1] = addc(Y[N - 1], 0u, cf);
    Y[0] = mad_lo_cc(m, P::mod(1), Y[0], cf);
    Y[1] = madc_hi_cc(m, P::mod(1), Y[1], cf);
#pragma unroll
    for (int j = 3; j < N; j += 2) {
        Y[j - 1] = madc_lo_cc(m, P::mod(j), Y[j - 1], cf);
        Y[j] = madc_hi_cc(m, P::mod(j), Y[j], cf);
    }
}
template <class P, int J>
struct SqrRows {
    static HD void run(uint32_t* X, uint32_t* Y, const uint32_t* a, const uint32_t* d) {
        mont_row_sq<P, J>(X, Y, a, d);
        SqrRows<P, J + 1>::run(Y, X, a, d);      // the two arrays swap roles every row
    }
};
template <class P>
struct SqrRows<P, P::N> {
    static HD void run(uint32_t*, uint32_t*, const uint32_t*, const uint32_t*) {}
};

template <class P>
HD Fe<P> fe_sqr(const Fe<P>& a) {
    constexpr int N = P::N;
    if (!P::FAST_SQR) return fe_mul(a, a);
    uint32_t ev[N], od[N], d[N];
    d[0] = a.l[0] << 1;
#pragma unroll
    for (int i = 1; i < N; i++) d[i] = (a.l[i] << 1) | (a.l[i - 1] >> 31);
    SqrRows<P, 0>::run(ev, od, a.l, d);
    // N rows (even count): the last row had X = od, as in fe_mul
    Fe<P> r; uint32_t cf = 0;
    r.l[0] = add_cc(od[1], ev[0], cf);
#pragma unroll
    for (int k = 1; k < N - 1; k++) r.l[k] = addc_cc(od[k + 1], ev[k], cf);
    r.l[N - 1] = addc(ev[N - 1], 0u, cf);
    fe_reduce_once(r);
    return r;
}

// canonical <-> Montgomery
template <class P>
HD Fe<P> fe_to_mont(const Fe<P>& a) { return fe_mul(a, Fe<P>::r2()); }
template <class P>
HD Fe<P> fe_from_mont(const Fe<P>& a) {
    Fe<P> o = Fe<P>::zero(); o.l[0] = 1;
    return fe_mul(a, o);
}

// a^e, e given as NE 32-bit limbs (not constant time; exponents here are public)
template <class P, int NE>
HD Fe<P> fe_pow(const Fe<P>& a, const uint32_t (&e)[NE]) {
    Fe<P> acc = Fe<P>::one();
    bool started = false;
    for (int i = NE * 32 - 1; i >= 0; i--) {
        if (started) acc = fe_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) { acc = started ? fe_mul(acc, a) : a; started = true; }
    }
    return acc;
}
// Fermat inverse (0 -> 0, as kilic's RedInverse does for bls/bignum_kilic.go:113)
template <class P>
HD Fe<P> fe_inv(const Fe<P>& a) {
    uint32_t e[P::N];
    for (int i = 0; i < P::N; i++) e[i] = P::modm2(i);
    if (a.is_zero()) return a;
    return fe_pow<P, P::N>(a, e);
}

typedef Fe<FpParams> Fp;
typedef Fe<FrParams> Fr;

}  // namespace b200
#include "field_fp64_impl.cuh"
#include "field_karatsuba.cuh"
#include "field_hybrid.cuh"
namespace b200 {

// ---------------------------------------------------------------------------------------
// Fp products as used by the curve code.  On the device they are calls to ONE out-of-line copy
// of the 381-bit Montgomery product (~350 SASS instructions) instead of an inline expansion per
// call site: a point addition with 16 inlined products is ~90 KB of straight-line code, the G1
// kernels then stall ~80 % of the time on instruction fetch (ncu: stall_no_inst) and ptxas,
// short of aligned register pairs, breaks IMAD.WIDE into IMAD + IMAD.HI.  Operands travel
// through local memory (L1-resident), 128-bit accesses.
// The out-of-line functions RETURN their result by value (PTX .param space, i.e. registers): an
// earlier form `Fp r; fp_mul_out(&r, &a, &b); return r;` left one temporary alloca with inliner
// lifetime markers per call site; nvcc 12.9 then forwarded a named variable onto such a
// temporary (`Fp c2 = fp_sqr(z)` became the temporary itself) without extending its lifetime, and
// stack colouring handed the slot to the next product's temporary while c2 was still live.
//
// Operands are copied into locals first: multiplying straight out of *a / *b made ptxas split every
// a_j * b_i product into IMAD + IMAD.HI + 2 IADD3.X instead of one IMAD.WIDE.X (ncu: half of the
// multiplier issue slots of the G1 kernels), with the copies all 288 products fuse.
//
// Two pipes (B200_FP64_SLOTS, default 0 = off): the integer product is bound by the IMAD.WIDE
// issue rate (1 per 4 cycles per SM sub-partition) while the FP64 pipe idles, so the warps of
// every other resident CTA (hardware warp slot / 4 = index of the CTA on its SM, CTAs being 4
// warps) run the bit-identical FP64-pipe product of field_fp64.cuh instead.
#ifdef __CUDA_ARCH__
#ifndef B200_FP64_SLOTS
#define B200_FP64_SLOTS 0
#endif
#if B200_FP64_SLOTS
__device__ __forceinline__ bool fp_on_fp64_pipe() {
    unsigned w; asm("mov.u32 %0, %%warpid;" : "=r"(w));
    return ((w >> 2) & 3u) < (unsigned)B200_FP64_SLOTS;
}
static __device__ __noinline__ Fp fp_mul_out(const Fp* a, const Fp* b) {
    Fp x = *a, y = *b;
    if (fp_on_fp64_pipe()) return fe_mul_fp64(x, y);
    return fe_mul(x, y);
}
static __device__ __noinline__ Fp fp_sqr_out(const Fp* a) {
    Fp x = *a;
    if (fp_on_fp64_pipe()) return fe_sqr_fp64(x);
    return fe_sqr(x);
}
#elif defined(B200_HYBRID_MUL)
// product half on the FP64 pipe, reduction half on the FMA-heavy pipe (field_hybrid.cuh); B200_HYBRID_MUL = 1: both, 2: product only
static __device__ __noinline__ Fp fp_mul_out(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_hyb(x, y); }
#if B200_HYBRID_MUL == 1
static __device__ __noinline__ Fp fp_sqr_out(const Fp* a) { Fp x = *a; return fe_sqr_hyb(x); }
#else
static __device__ __noinline__ Fp fp_sqr_out(const Fp* a) { Fp x = *a; return fe_sqr(x); }
#endif
#elif defined(B200_KARATSUBA)
// Karatsuba product half + m * p-only reduction rows (field_karatsuba.cuh); the square keeps fe_sqr
// (the Karatsuba square needs more FMA-pipe cycles than it saves once its carry adds are counted).
static __device__ __noinline__ Fp fp_mul_out(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_k(x, y); }
static __device__ __noinline__ Fp fp_sqr_out(const Fp* a) { Fp x = *a; return fe_sqr(x); }
#else
static __device__ __noinline__ Fp fp_mul_out(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul(x, y); }
static __device__ __noinline__ Fp fp_sqr_out(const Fp* a) { Fp x = *a; return fe_sqr(x); }
#endif
__device__ __forceinline__ Fp fp_mul(const Fp& a, const Fp& b) { return fp_mul_out(&a, &b); }
__device__ __forceinline__ Fp fp_sqr(const Fp& a) { return fp_sqr_out(&a); }
#else
inline Fp fp_mul(const Fp& a, const Fp& b) { return fe_mul(a, b); }
inline Fp fp_sqr(const Fp& a) { return fe_sqr(a); }
#endif

#ifdef __CUDACC__
// Whole-element loads / stores.  Fe (and everything built from it) is alignas(16), so a plain
// struct copy lowers to 128-bit vector accesses (LDG.128 / STS.128 / ...).  Do NOT reinterpret
// elements as uint4: that breaks the aliasing rules and nvcc then reorders / drops the copies
// (observed: the first 16 bytes of a point silently not copied).
template <class T>
__device__ __forceinline__ T ld_vec(const T* p) {
    static_assert(alignof(T) >= 16 && sizeof(T) % 16 == 0, "16 B aligned element");
    return *p;
}
template <class T>
__device__ __forceinline__ void st_vec(T* p, const T& v) {
    static_assert(alignof(T) >= 16 && sizeof(T) % 16 == 0, "16 B aligned element");
    *p = v;
}
#endif

}  // namespace b200
