#include <stdio.h>
#include <stdlib.h>
#include "g1_dev.cuh"
#include "hostutil.cuh"
using namespace b200;
static uint64_t st = 88172645463325252ULL;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); }
int main() {
    int bad = 0;
    for (int it = 0; it < 40; it++) {
        Fr k, b;
        for (int i = 0; i < 8; i++) { k.l[i] = rnd(); b.l[i] = rnd(); }
        k.l[7] &= 0x3fffffffu; b.l[7] &= 0x3fffffffu;
        if (it == 0) { k = Fr::zero(); k.l[0] = 1; }
        if (it == 1) { k = Fr::zero(); k.l[0] = 2; }
        if (it == 2) { k = Fr::zero(); k.l[0] = 17; }
        if (it == 3) { for (int i = 0; i < 8; i++) k.l[i] = FrParams::mod(i); k.l[0] -= 1; }
        G1J g = g1_generator();
        G1J base = it % 7 == 6 ? g : g1_mul_simple(g, b.l);
        G1J want = g1_mul_simple(base, k.l);
        for (int mode = 0; mode < 2; mode++) {
            ScalarProgram sp;
            make_scalar_program(&sp, k, mode);
            G1J got;
            g1_mul_digits(&got, &base, sp.d1, sp.d2, sp.top, sp.mode);
            if (!g1_equal(got, want)) { bad++; printf("mismatch it=%d mode=%d\n", it, mode); }
            G1J got2;
            g1_mul_digits_jac(&got2, &base, sp.d1, sp.d2, sp.top, sp.mode);
            if (!g1_equal(got2, want)) { bad++; printf("mismatch(jac) it=%d mode=%d\n", it, mode); }
        }
        G1J got;
        g1_mul_var(&got, &base, k.l);
        if (!g1_equal(got, want)) { bad++; printf("mismatch var it=%d\n", it); }
    }
    printf("bad=%d\n", bad);
    return bad != 0;
}
