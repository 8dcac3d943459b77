#include <stdio.h>
#include "field.cuh"
using namespace b200;
static uint64_t st = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
static Fp rnd_fp(int kind) {
    Fp a;
    for (int i = 0; i < 12; i++) a.l[i] = rnd();
    a.l[11] &= 0x0fffffffu;
    if (kind == 1) a = Fp::zero();
    if (kind == 2) { a = Fp::zero(); a.l[0] = 1; }
    if (kind == 3) { a = Fp::modulus(); a.l[0] -= 1; }
    if (kind == 4) { for (int i = 0; i < 12; i++) a.l[i] = 0xffffffffu; a.l[11] = 0x0fffffffu; }
    if (kind == 5) { a = Fp::modulus(); a.l[0] -= 2; }
    if (kind == 6) { for (int i = 0; i < 6; i++) a.l[i] = 0xffffffffu; for (int i = 6; i < 12; i++) a.l[i] = 0; }
    if (kind == 7) { for (int i = 0; i < 6; i++) a.l[i] = 0; for (int i = 6; i < 12; i++) a.l[i] = 0xffffffffu; a.l[11] = 0x0fffffffu; }
    if (kind == 8) { for (int i = 0; i < 12; i++) a.l[i] = (i < 6) ? 5u : 5u; }       /* a_lo == a_hi */
    return a;
}
int main() {
    int bad = 0;
    for (int it = 0; it < 20000; it++) {
        Fp a = rnd_fp(it < 81 ? it % 9 : 0), b = rnd_fp(it < 81 ? it / 9 : 0);
        if (fe_mul(a, b) != fe_mul_k(a, b)) { bad++; if (bad < 5) printf("mul mismatch it=%d\n", it); }
        if (fe_mul_portable(a, a) != fe_sqr_k(a)) { bad++; if (bad < 5) printf("sqr mismatch it=%d\n", it); }
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
