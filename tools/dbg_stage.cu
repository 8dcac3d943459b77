// debug: one G1 FFT on device vs the same stage logic on the host
#include "../go_kzg_b200/csrc/kernels_g1.cu"
#include "../go_kzg_b200/csrc/hostutil.cuh"
#include <cstdio>
#include <vector>
namespace b200 { bool g_prof_on = false; void prof_begin_event(int, cudaStream_t) {} void prof_end_event(cudaStream_t) {} }
using namespace b200;
template <int V>
__global__ void k_variant(G1J* data, size_t m, const ScalarProgram* progs, size_t prog_stride) {
    size_t q = threadIdx.x;
    size_t j = q & (m - 1);
    size_t i0 = 2 * q - j, i1 = i0 + m;
    G1J* p0 = data + i0; G1J* p1 = data + i1;
    const ScalarProgram* prog = progs + j * prog_stride;
    G1J x0 = ld_vec(p0), x1 = ld_vec(p1), s, d, r;
    if (V == 0) { g1_add_sub_ni(&s, &d, &x0, &x1); st_vec(p0, s); st_vec(p1, d); }
    if (V == 1) { g1_add_sub_ni(&s, &d, &x0, &x1); g1_mul_program(&r, &d, prog); st_vec(p0, s); st_vec(p1, r); }
    if (V == 2) { s = g1_add(x0, x1); d = g1_sub(x0, x1); g1_mul_program(&r, &d, prog); st_vec(p0, s); st_vec(p1, r); }
    if (V == 4) { g1_add_sub_ni(&s, &d, &x0, &x1); G1J d2 = d; g1_mul_program(&r, &d2, prog); st_vec(p0, s); st_vec(p1, r); }
    if (V == 5) { g1_add_ni(&s, &x0, &x1); G1J nx = g1_neg(x1); g1_add_ni(&d, &x0, &nx); g1_mul_program(&r, &d, prog); st_vec(p0, s); st_vec(p1, r); }
    if (V == 6) { g1_add_sub_ni(&s, &d, &x0, &x1); if (prog->is_one) r = d; else g1_mul_digits(&r, &d, prog->d1, prog->d2, prog->top, 0); st_vec(p0, s); st_vec(p1, r); }
    if (V == 7) { g1_add_sub_ni(&s, &d, &x0, &x1); Fr k = Fr::zero(); k.l[0] = 12345 + (uint32_t)q; g1_mul_var(&r, &d, k.l); st_vec(p0, s); st_vec(p1, r); }
    if (V == 3) { g1_add_sub_ni(&s, &d, &x0, &x1); st_vec(p1, d); __syncthreads(); d = ld_vec(p1); g1_mul_program(&r, &d, prog); st_vec(p0, s); st_vec(p1, r); }
}
int main() {
    const unsigned scale = 3; const size_t W = 8, n = 8;
    Fr w = fe_to_mont(fr_scale2_root_canon(scale));
    std::vector<Fr> ex(W + 1); ex[0] = Fr::one(); for (size_t i = 1; i <= W; i++) ex[i] = fe_mul(ex[i-1], w);
    std::vector<ScalarProgram> progs(W / 2);
    for (size_t j = 0; j < W / 2; j++) make_scalar_program(&progs[j], fe_from_mont(ex[j]), 0);
    G1J g = g1_generator();
    std::vector<G1J> x(n), data(n), dev(n);
    for (size_t i = 0; i < n; i++) { uint32_t k[8] = {(uint32_t)(i + 1),0,0,0,0,0,0,0}; x[i] = g1_mul_simple(g, k); data[i] = x[i]; }
    G1J* d_data; ScalarProgram* d_progs;
    cudaMalloc(&d_data, n * sizeof(G1J)); cudaMalloc(&d_progs, progs.size() * sizeof(ScalarProgram));
    cudaMemcpy(d_data, x.data(), n * sizeof(G1J), cudaMemcpyHostToDevice);
    cudaMemcpy(d_progs, progs.data(), progs.size() * sizeof(ScalarProgram), cudaMemcpyHostToDevice);
    size_t halfw = W / 2;
    for (size_t m = n / 2; m >= 1; m >>= 1) {
        for (size_t q = 0; q < n / 2; q++) {
            size_t j = q & (m - 1), i0 = 2 * q - j, i1 = i0 + m;
            G1J s, d, t;
            g1_add_sub_ni(&s, &d, &data[i0], &data[i1]);
            g1_mul_program(&t, &d, &progs[j * (halfw / m)]);
            data[i0] = s; data[i1] = t;
        }
        launch_g1_fft_stage(d_data, n / 2, 1, m, 1, n, true, d_progs, halfw / m, 0);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(dev.data(), d_data, n * sizeof(G1J), cudaMemcpyDeviceToHost);
        printf("stage m=%zu (%s):", m, cudaGetErrorString(e));
        for (size_t i = 0; i < n; i++) printf(" %s", g1_equal(dev[i], data[i]) ? "ok" : "BAD");
        printf("\n");
    }
    for (int v = 0; v < 8; v++) {
        const size_t m = 4;
        cudaMemcpy(d_data, x.data(), n * sizeof(G1J), cudaMemcpyHostToDevice);
        if (v == 0) k_variant<0><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 1) k_variant<1><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 2) k_variant<2><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 3) k_variant<3><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 4) k_variant<4><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 5) k_variant<5><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 6) k_variant<6><<<1, 4>>>(d_data, m, d_progs, 1);
        if (v == 7) k_variant<7><<<1, 4>>>(d_data, m, d_progs, 1);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(dev.data(), d_data, n * sizeof(G1J), cudaMemcpyDeviceToHost);
        printf("variant %d (%s):", v, cudaGetErrorString(e));
        for (size_t q = 0; q < 4; q++) {
            G1J s2, d2, t2;
            g1_add_sub_ni(&s2, &d2, &x[q], &x[q + 4]);
            g1_mul_program(&t2, &d2, &progs[q]);
            if (v == 7) { Fr k = Fr::zero(); k.l[0] = 12345 + (uint32_t)q; t2 = g1_mul_simple(d2, k.l); }
            printf(" [%s %s]", g1_equal(dev[q], s2) ? "ok" : "BAD", g1_equal(dev[q + 4], v == 0 ? d2 : t2) ? "ok" : "BAD");
        }
        printf("\n");
    }
    // per-program multiplication on device through launch_g1_mul_programs
    cudaMemcpy(d_data, x.data(), n * sizeof(G1J), cudaMemcpyHostToDevice);
    launch_g1_mul_programs(d_data, W / 2, 1, 1, n, d_progs, 1, 0, 0, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(dev.data(), d_data, n * sizeof(G1J), cudaMemcpyDeviceToHost);
    printf("mul_programs:");
    for (size_t i = 0; i < W / 2; i++) { G1J t; g1_mul_program(&t, &x[i], &progs[i]); printf(" %s", g1_equal(dev[i], t) ? "ok" : "BAD"); }
    printf("\n");
    return 0;
}
