// CPU replay of the Pippenger pipeline of msm.cuh (the kernels of kernels_msm.cu run these bodies one thread each)
// against plain double-and-add: random and edge scalars, points at infinity, repeated points (doubling inside
// buckets), P and -P in one bucket (cancellation), affine and Jacobian inputs.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "msm.cuh"
#include "hostutil.cuh"
using namespace b200;
static uint64_t st = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }

static G1J msm_replay(const std::vector<G1J>& pts, const std::vector<Fr>& ks) {
    const size_t n = pts.size();
    MsmPlan p = msm_plan(n);
    const size_t slots = (size_t)p.W * (p.B + 1);
    std::vector<int16_t> digits((size_t)p.W * p.T);
    std::vector<uint32_t> counts(slots, 0), offsets(slots, 0), cursors(slots, 0), sorted((size_t)p.W * p.T, 0);
    std::vector<Fp> bx(n);
    uint32_t not_affine = 0;
    for (size_t i = 0; i < n; i++) msm_recode_point(p, i, pts[i], ks[i], digits.data(), counts.data(), bx.data(), &not_affine);
    for (unsigned w = 0; w < p.W; w++) {
        uint32_t run = 0;
        for (unsigned b = 0; b <= p.B; b++) { offsets[(size_t)w * (p.B + 1) + b] = run; run += counts[(size_t)w * (p.B + 1) + b]; }
    }
    for (unsigned w = 0; w < p.W; w++)
        for (size_t t = 0; t < p.T; t++) msm_scatter_term(p, w, t, digits.data(), offsets.data(), cursors.data(), sorted.data());
    std::vector<G1J> buckets((size_t)p.W * p.B, G1J::infinity());
    // units exactly as k_msm_accumulate maps them: window w has Bw[w] buckets x Sw[w] slices
    for (unsigned w = 0; w < p.W; w++) {
        for (unsigned b = 1; b <= p.B; b++) {
            const size_t slot = (size_t)w * (p.B + 1) + b;
            if (b > p.Bw[w]) { if (counts[slot]) { printf("digit beyond Bw in window %u\n", w); exit(2); } continue; }
            const unsigned S = p.Sw[w];
            const uint32_t len = counts[slot], off = offsets[slot], per = (len + S - 1) / S;
            G1J sum = G1J::infinity();
            for (unsigned sl = 0; sl < S; sl++) {
                uint32_t begin = sl * per, end = begin + per;
                if (begin > len) begin = len;
                if (end > len) end = len;
                G1J part;
                msm_accumulate_slice(p, pts.data(), bx.data(), sorted.data() + (size_t)w * p.T, off + begin, off + end, not_affine == 0, &part);
                sum = g1_add(sum, part);
            }
            buckets[(size_t)w * p.B + (b - 1)] = sum;
        }
    }
    G1J total = G1J::infinity();
    const unsigned nseg = p.B / p.L;
    for (unsigned w = 0; w < p.W; w++) {
        G1J ws = G1J::infinity();
        for (unsigned s = 0; s < nseg; s++) {
            G1J v;
            msm_reduce_segment(buckets.data() + (size_t)w * p.B, s * p.L + 1, p.L, &v);
            ws = g1_add(ws, v);
        }
        for (unsigned d = 0; d < p.c * w; d++) ws = g1_dbl(ws);
        total = g1_add(total, ws);
    }
    return total;
}

int main(int argc, char** argv) {
    int bad = 0;
    const size_t sizes[] = {32, 33, 100, 700};
    for (size_t n : sizes) {
        for (int variant = 0; variant < 2; variant++) {
            std::vector<G1J> pts(n);
            std::vector<Fr> ks(n);
            G1J g = g1_generator();
            G1J want = G1J::infinity();
            for (size_t i = 0; i < n; i++) {
                Fr b, k;
                for (int j = 0; j < 8; j++) { b.l[j] = rnd(); k.l[j] = rnd(); }
                b.l[7] &= 0x3fffffffu; k.l[7] &= 0x3fffffffu;
                for (int j = 2; j < 8; j++) b.l[j] = 0;                       // cheap base points: 64-bit multiples of G
                G1J P = g1_mul_simple(g, b.l);
                if (variant == 0) {                                             // affine inputs (Z = 1)
                    Fp zi = fe_inv(P.z), zi2 = fe_sqr(zi);
                    P.x = fe_mul(P.x, zi2); P.y = fe_mul(P.y, fe_mul(zi2, zi)); P.z = Fp::one();
                }
                if (i == 1) k = Fr::zero();
                if (i == 2) { k = Fr::zero(); k.l[0] = 1; }
                if (i == 3) { for (int j = 0; j < 8; j++) k.l[j] = FrParams::mod(j); k.l[0] -= 1; }   // r - 1
                if (i == 4) P = G1J::infinity();
                if (i == 6) { P = pts[5]; k = ks[5]; }                          // same term twice: doubling inside a bucket
                if (i == 8) { P = g1_neg(pts[7]); k = ks[7]; }                  // P and -P with one scalar: cancellation
                if (i == 9) { k = Fr::zero(); k.l[3] = 0x80000000u; }           // 2^127
                pts[i] = P; ks[i] = k;
                want = g1_add(want, g1_mul_simple(P, k.l));
            }
            G1J got = msm_replay(pts, ks);
            if (!g1_equal(got, want)) { bad++; printf("mismatch n=%zu variant=%d\n", n, variant); }
            MsmPlan p = msm_plan(n);
            printf("n=%zu variant=%d c=%u W=%u B=%u S=%u L=%u top: Bw=%u Sw=%u units=%u\n", n, variant, p.c, p.W, p.B, p.S, p.L, p.Bw[p.W - 1], p.Sw[p.W - 1], p.unit_off[p.W]);
        }
    }
    for (size_t n : {(size_t)4096, (size_t)65536, (size_t)1 << 20}) {
        MsmPlan p = msm_plan(n);
        printf("plan n=%zu: c=%u W=%u B=%u S=%u L=%u\n", n, p.c, p.W, p.B, p.S, p.L);
        for (unsigned w = 0; w < p.W; w++) printf("   window %u: Bw=%u Sw=%u units at %u\n", w, p.Bw[w], p.Sw[w], p.unit_off[w]);
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
