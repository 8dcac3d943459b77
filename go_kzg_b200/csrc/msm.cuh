// msm.cuh -- Pippenger bucket method for variable-base multi-scalar multiplication
// (bls.LinCombG1, bls/bls_kilic.go:132-150 -> kilic MultiExp), per-thread bodies.
//
// sum_i k_i P_i for arbitrary points.  Every scalar is split with the GLV endomorphism into two
// 128-bit halves k = k1 + k2 z^2 (z^2 (x, y) = (beta x, -y)), which doubles the term count and halves
// the window count; each half is cut into W = ceil(129 / c) signed c-bit digits d in [-2^(c-1), 2^(c-1)]:
//     sum_i k_i P_i = sum_w 2^(c w) sum_b b * Bucket[w][b],   Bucket[w][b] = sum of the +-terms with |digit| = b.
// Pipeline (kernels_msm.cu launches one kernel per step; every body below is __host__ __device__ so that
// tools/msm_host_check.cu replays the whole algorithm on the CPU against double-and-add):
//   1 recode    per point: GLV split, digits of both halves, digit histogram per window, beta x
//   2 scan      per window: exclusive prefix of the histogram -> start of every bucket's term list
//   3 scatter   counting sort of the terms by (window, bucket); sign in bit 31
//   4 accumulate  unit = (bucket, slice): sums its share of the bucket's list; the S slices of a bucket sit
//               in adjacent lanes and are combined with a warp-shuffle tree
//   5 segments  unit = L consecutive buckets: sum_j (b0 + j) Bucket[b0 + j] with a running sum
//   6 windows   one CTA per window: tree sum of its segments
//     (windows of up to 1024 buckets: 5 + 6 are one kernel, k_msm_window_scan(_cluster) -- running sums per quad, suffix scan
//      over the quads (of a cluster of four CTAs from 128 buckets on), one tree; no small multiplication by a base index)
//   7 horner    window w doubled c w times, then the sum of the W weighted window sums
// Steps 5-7 (and step 4 when the problem is too small to fill the chip) run one unit per QUAD of lanes (quad.cuh):
// they are waited for because of the depth of their chains of point operations, and a quad walks a chain 2-3x faster.
// Terms: t < n is (k1 digit, P_t), t >= n is (k2 digit, z^2 P_(t - n)).
#pragma once
#include "g1_dev.cuh"

namespace b200 {

#define MSM_MAX_WINDOWS 34
struct MsmPlan {
    size_t n = 0;       // points
    size_t T = 0;       // terms = 2 n
    unsigned c = 0;     // window bits
    unsigned W = 0;     // windows
    unsigned B = 0;     // buckets per window = 2^(c-1), bucket index 1..B
    unsigned S = 1;     // base number of slices per bucket in the accumulation (power of two <= 32)
    unsigned L = 1;     // buckets per segment in the reduction (power of two)
    // Accumulation units.  The windows are not equally loaded: the top window holds only the 128 - c (W - 1) leading bits of
    // a half scalar, so its digits are non-negative and below 2^bits -- fewer buckets, each proportionally longer -- and a
    // last window may hold nothing but the carry.  Window w gets Bw[w] buckets x Sw[w] slices with Bw Sw (about) constant,
    // so that every (bucket, slice) unit adds the same number of terms; units of one window are padded to whole warps.
    unsigned Bw[MSM_MAX_WINDOWS] = {};
    unsigned Sw[MSM_MAX_WINDOWS] = {};
    unsigned unit_off[MSM_MAX_WINDOWS + 1] = {};
};

// thread slots of one full wave of the accumulation kernel (148 SMs x 4 CTAs x 128 threads at 124 registers)
static const size_t kMsmWaveSlots = (size_t)148 * 4 * 128;

// Window width minimising (additions into buckets) + (additions of the bucket reduction); slices so that the
// accumulation has as many units as fit ONE wave of the chip (a second, nearly empty wave would double its time)
// while each slice keeps a few terms.
inline MsmPlan msm_plan(size_t n) {
    MsmPlan p;
    p.n = n; p.T = 2 * n;
    double best = 0;
    for (unsigned c = 4; c <= 14; c++) {
        unsigned W = (129 + c - 1) / c;
        double cost = (double)p.T * W + 2.0 * W * (double)(1u << (c - 1));
        if (p.c == 0 || cost < best) { best = cost; p.c = c; }
    }
    p.W = (129 + p.c - 1) / p.c;
    p.B = 1u << (p.c - 1);
    const double avg = (double)p.T / p.B;                       // terms per bucket (upper bound: zero digits drop out)
    // Slices per bucket.  Measured (profiles/r02_msm_accumulate_2p20.txt): a warp of this kernel advances at its own
    // latency-bound pace (~60 k cycles per mixed addition) whether 1 or 4 warps share a sub-partition, so the time is
    // (rounds of resident warps) x (chain of one unit): minimise ceil(units / slots) x (avg / S additions + log2 S tree levels).
    p.S = 1;
    double best_t = 0;
    for (unsigned S = 1, lg = 0; S <= 32 && (S == 1 || avg / S >= 2.0); S *= 2, lg++) {
        const size_t units = (size_t)p.W * p.B * S;
        const double rounds = (double)((units + kMsmWaveSlots - 1) / kMsmWaveSlots);
        const double t = rounds * (avg / S * 11.0 + lg * 16.0);
        if (S == 1 || t < best_t) { best_t = t; p.S = S; }
    }
    p.L = 4;
    if (p.L > p.B) p.L = p.B;
    unsigned off = 0;
    for (unsigned w = 0; w < p.W; w++) {
        const int bits = 128 - (int)(p.c * w);                  // bits of a 128-bit half scalar that reach this window
        unsigned bw = p.B;
        if (bits < (int)p.c) bw = bits <= 0 ? 1u : (1u << bits);   // digit = bits + carry <= 2^bits: never negative there
        if (bw > p.B) bw = p.B;
        unsigned sw = p.S * (p.B / bw);
        if (sw > 32) sw = 32;
        p.Bw[w] = bw; p.Sw[w] = sw;
        p.unit_off[w] = off;
        off += (bw * sw + 31u) / 32u * 32u;
    }
    p.unit_off[p.W] = off;
    return p;
}

// k (canonical, < r) = k1 + k2 z^2 with k1 < z^2 < 2^128, k2 < 2^128: long division by the 128-bit constant z^2
__host__ __device__ inline void msm_glv_split(const uint32_t k[8], uint32_t k1[4], uint32_t k2[4]) {
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
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = quo[i]; }
}

// signed c-bit digits of a 128-bit value: v = sum_w out[w * stride] 2^(c w), |digit| <= 2^(c-1)
__host__ __device__ inline void msm_digits(const uint32_t v[4], unsigned c, unsigned W, int16_t* out, size_t stride) {
    const unsigned half = 1u << (c - 1), mask = (1u << c) - 1u;
    unsigned carry = 0;
    for (unsigned w = 0; w < W; w++) {
        const unsigned off = w * c;
        unsigned bits = 0;
        if (off < 128) {
            const unsigned li = off >> 5, sh = off & 31u;
            uint64_t two = v[li];
            if (li + 1 < 4) two |= (uint64_t)v[li + 1] << 32;
            bits = (unsigned)(two >> sh) & mask;
        }
        unsigned d = bits + carry;
        int dd;
        if (d > half) { dd = (int)d - (int)(mask + 1u); carry = 1; } else { dd = (int)d; carry = 0; }
        out[(size_t)w * stride] = (int16_t)dd;
    }
}

__host__ __device__ inline uint32_t msm_atomic_inc(uint32_t* p) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}

// step 1, point i.  digits: [W][T] int16; counts: [W][B + 1] (index = |digit|); bx[i] = beta x_i;
// *not_affine is raised when a finite point has Z != 1 (the accumulation then uses general additions).
__host__ __device__ inline void msm_recode_point(const MsmPlan& p, size_t i, const G1J& pt, const Fr& k_canon, int16_t* digits,
                                                 uint32_t* counts, Fp* bx, uint32_t* not_affine) {
    uint32_t k1[4], k2[4];
    msm_glv_split(k_canon.l, k1, k2);
    const bool inf = pt.is_inf();
    if (inf) { for (int j = 0; j < 4; j++) { k1[j] = 0; k2[j] = 0; } }       // contributes nothing
    msm_digits(k1, p.c, p.W, digits + i, p.T);
    msm_digits(k2, p.c, p.W, digits + p.n + i, p.T);
    for (unsigned w = 0; w < p.W; w++) {
        int d1 = digits[(size_t)w * p.T + i], d2 = digits[(size_t)w * p.T + p.n + i];
        if (d1) msm_atomic_inc(counts + (size_t)w * (p.B + 1) + (d1 < 0 ? -d1 : d1));
        if (d2) msm_atomic_inc(counts + (size_t)w * (p.B + 1) + (d2 < 0 ? -d2 : d2));
    }
    bx[i] = fp_mul(pt.x, fp_const_beta());
    if (!inf && pt.z != Fp::one()) *not_affine = 1u;
}

// step 3, (window w, term t): sorted[w][offsets[w][b] + k] = t | sign << 31
__host__ __device__ inline void msm_scatter_term(const MsmPlan& p, unsigned w, size_t t, const int16_t* digits, const uint32_t* offsets,
                                                 uint32_t* cursors, uint32_t* sorted) {
    int d = digits[(size_t)w * p.T + t];
    if (!d) return;
    const unsigned b = (unsigned)(d < 0 ? -d : d);
    const size_t slot = (size_t)w * (p.B + 1) + b;
    uint32_t pos = offsets[slot] + msm_atomic_inc(cursors + slot);
    sorted[(size_t)w * p.T + pos] = (uint32_t)t | (d < 0 ? 0x80000000u : 0u);
}

// step 4, one slice of one bucket: sum of the terms sorted_w[begin .. end)
__host__ __device__ inline void msm_accumulate_slice(const MsmPlan& p, const G1J* pts, const Fp* bx, const uint32_t* sorted_w, uint32_t begin,
                                                     uint32_t end, bool affine, G1J* acc_out) {
    G1J acc = G1J::infinity();
    for (uint32_t e = begin; e < end; e++) {
        const uint32_t u = sorted_w[e];
        const size_t t = u & 0x7fffffffu;
        const bool second = t >= p.n;                               // k2 term: z^2 P = (beta x, -y)
        const size_t i = second ? t - p.n : t;
        const bool neg = ((u >> 31) != 0) != second;
        if (affine) {
            G1A q;
            q.x = second ? bx[i] : pts[i].x;
            q.y = pts[i].y;
            if (neg) q.y = fe_neg(q.y);
            g1_add_mixed_ni(&acc, &acc, &q);
        } else {
            G1J q;
            q.x = second ? bx[i] : pts[i].x;
            q.y = pts[i].y;
            q.z = pts[i].z;
            if (neg) q.y = fe_neg(q.y);
            g1_add_ni(&acc, &acc, &q);
        }
    }
    *acc_out = acc;
}

// m * P for a small m (bucket indices: at most 13 bits)
__host__ __device__ inline void msm_small_mul(G1J* out, const G1J* p, unsigned m) {
    G1J acc = G1J::infinity();
    for (int bit = 15; bit >= 0; bit--) {
        if (!acc.is_inf()) g1_dbl_ni(&acc, &acc);
        if ((m >> bit) & 1u) g1_add_ni(&acc, &acc, p);
    }
    *out = acc;
}

// step 5: sum_(j < L) (b0 + j) * bkt[b0 - 1 + j]   (bkt indexed by bucket - 1; b0 >= 1)
__host__ __device__ inline void msm_reduce_segment(const G1J* bkt, unsigned b0, unsigned L, G1J* out) {
    G1J running = G1J::infinity(), acc = G1J::infinity();
    for (int j = (int)L - 1; j >= 0; j--) {
        g1_add_ni(&running, &running, bkt + (b0 - 1 + j));
        g1_add_ni(&acc, &acc, &running);
    }
    // acc = sum (j + 1) bkt[..]; the remaining (b0 - 1) * (sum of the segment) by double-and-add
    if (b0 > 1 && !running.is_inf()) {
        G1J m;
        msm_small_mul(&m, &running, b0 - 1);
        g1_add_ni(&acc, &acc, &m);
    }
    *out = acc;
}

}  // namespace b200
