#include <stdio.h>
#include "field_fp64.cuh"
using namespace b200;
static uint64_t st = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
static Fp rnd_fp(int kind) {
    Fp a;
    for (int i = 0; i < 12; i++) a.l[i] = rnd();
    a.l[11] &= 0x0fffffffu;   // < 2^380 < p
    if (kind == 1) a = Fp::zero();
    if (kind == 2) { a = Fp::zero(); a.l[0] = 1; }
    if (kind == 3) { a = Fp::modulus(); a.l[0] -= 1; }        // p - 1
    if (kind == 4) { for (int i = 0; i < 12; i++) a.l[i] = 0xffffffffu; a.l[11] = 0x0fffffffu; }
    if (kind == 5) { a = Fp::modulus(); a.l[0] -= 2; }
    return a;
}
int main() {
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        Fp a = rnd_fp(it < 36 ? it % 6 : 0), b = rnd_fp(it < 36 ? it / 6 : 0);
        Fp w = fe_mul(a, b), g = fe_mul_fp64(a, b);
        if (w != g) { bad++; if (bad < 5) printf("mul mismatch it=%d\n", it); }
        Fp h3m = fe_mul_hyb3(a, b), h3s = fe_sqr_hyb3(a);
        if (w != h3m) { bad++; if (bad < 5) printf("hybrid3 mul mismatch it=%d\n", it); }
        if (fe_sqr(a) != h3s) { bad++; if (bad < 5) printf("hybrid3 sqr mismatch it=%d\n", it); }
        Fp h2m = fe_mul_hyb2(a, b), h2s = fe_sqr_hyb2(a);
        if (w != h2m) { bad++; if (bad < 5) printf("hybrid2 mul mismatch it=%d\n", it); }
        if (fe_sqr(a) != h2s) { bad++; if (bad < 5) printf("hybrid2 sqr mismatch it=%d\n", it); }
        Fp hm = fe_mul_hyb(a, b), hs = fe_sqr_hyb(a);
        if (w != hm) { bad++; if (bad < 5) printf("hybrid mul mismatch it=%d\n", it); }
        if (fe_sqr(a) != hs) { bad++; if (bad < 5) printf("hybrid sqr mismatch it=%d\n", it); }
        Fp ws = fe_sqr(a), gs = fe_sqr_fp64(a);
        if (ws != gs) { bad++; if (bad < 5) printf("sqr mismatch it=%d\n", it); }
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
