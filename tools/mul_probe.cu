// mul_probe.cu -- throughput of the 381-bit Montgomery product on the integer pipe (fe_mul), on the
// FP64 pipe (fe_mul_fp64) and with the warps of every SM sub-partition split between the two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I go_kzg_b200/csrc -I include -o tools/mul_probe tools/mul_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field_fp64.cuh"
using namespace b200;

static __device__ __noinline__ Fp mul_int(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul(x, y); }
static __device__ __noinline__ Fp mul_fp(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_fp64(x, y); }
static __device__ __noinline__ Fp sqr_int(const Fp* a) { Fp x = *a; return fe_sqr(x); }
static __device__ __noinline__ Fp sqr_fp(const Fp* a) { Fp x = *a; return fe_sqr_fp64(x); }
static __device__ __noinline__ Fp mul_hy(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_hyb(x, y); }
static __device__ __noinline__ Fp sqr_hy(const Fp* a) { Fp x = *a; return fe_sqr_hyb(x); }
static __device__ __noinline__ Fp mul_h3(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_hyb3(x, y); }
static __device__ __noinline__ Fp sqr_h3(const Fp* a) { Fp x = *a; return fe_sqr_hyb3(x); }
static __device__ __noinline__ Fp mul_h2(const Fp* a, const Fp* b) { Fp x = *a, y = *b; return fe_mul_hyb2(x, y); }
static __device__ __noinline__ Fp sqr_h2(const Fp* a) { Fp x = *a; return fe_sqr_hyb2(x); }

__device__ __forceinline__ unsigned hw_warp_slot() { unsigned w; asm volatile("mov.u32 %0, %%warpid;" : "=r"(w)); return w; }

// mode: 0 integer only, 1 FP64 only, 2.. : FP64 when ((slot / 4) % mode_den) < mode_num
__global__ void __launch_bounds__(128, 4) k_probe(uint32_t* buf, int iters, int num, int den, int sqr, unsigned long long* check) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fp x, y;
    for (int k = 0; k < 12; k++) { x.l[k] = buf[k] + (uint32_t)i; y.l[k] = buf[12 + k] ^ (uint32_t)i; }
    x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu;
    bool fp = ((hw_warp_slot() >> 2) % den) < num;
    if (den == -4) {    // interleaved hybrid, instruction order pinned by data dependences
        if (sqr) for (int k = 0; k < iters; k++) { x = sqr_h3(&x); y = sqr_h3(&y); }
        else for (int k = 0; k < iters; k++) { x = mul_h3(&x, &y); y = mul_h3(&y, &x); }
    } else if (den == -3) {    // interleaved hybrid (rows of the reduction issued between the product columns)
        if (sqr) for (int k = 0; k < iters; k++) { x = sqr_h2(&x); y = sqr_h2(&y); }
        else for (int k = 0; k < iters; k++) { x = mul_h2(&x, &y); y = mul_h2(&y, &x); }
    } else if (den < 0) {      // hybrid product (FP64 product half, integer reduction half), every warp; den == -2: every other warp slot
        bool hy = den == -1 || (hw_warp_slot() & 1);
        if (sqr) { if (hy) for (int k = 0; k < iters; k++) { x = sqr_hy(&x); y = sqr_hy(&y); } else for (int k = 0; k < iters; k++) { x = sqr_int(&x); y = sqr_int(&y); } }
        else { if (hy) for (int k = 0; k < iters; k++) { x = mul_hy(&x, &y); y = mul_hy(&y, &x); } else for (int k = 0; k < iters; k++) { x = mul_int(&x, &y); y = mul_int(&y, &x); } }
    } else if (sqr) {
        if (fp) for (int k = 0; k < iters; k++) { x = sqr_fp(&x); y = sqr_fp(&y); }
        else for (int k = 0; k < iters; k++) { x = sqr_int(&x); y = sqr_int(&y); }
    } else {
        if (fp) for (int k = 0; k < iters; k++) { x = mul_fp(&x, &y); y = mul_fp(&y, &x); }
        else for (int k = 0; k < iters; k++) { x = mul_int(&x, &y); y = mul_int(&y, &x); }
    }
    unsigned long long acc = 0;
    for (int k = 0; k < 12; k++) acc = acc * 1000003ull + x.l[k] + 31ull * y.l[k];
    check[i] = acc;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    size_t threads = (size_t)sms * 4 * 128;     // one full wave at 4 CTAs per SM
    uint32_t h[24]; for (int i = 0; i < 24; i++) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    unsigned long long *c0, *c1; cudaMalloc(&c0, threads * 8); cudaMalloc(&c1, threads * 8);
    const int iters = 1000;
    struct { const char* name; int num, den; } modes[] = {{"integer pipe only", 0, 1}, {"FP64 pipe only", 1, 1}, {"2 of 4 warp slots on FP64", 1, 2},
                                                          {"hybrid (FP64 product + int reduction)", 1, -1}, {"hybrid on odd warp slots", 1, -2}, {"hybrid, interleaved rows", 1, -3}, {"hybrid, interleaved + pinned order", 1, -4}};
    unsigned long long* href = (unsigned long long*)malloc(threads * 8);
    unsigned long long* hgot = (unsigned long long*)malloc(threads * 8);
    for (int sqr = 0; sqr < 2; sqr++) {
        for (auto& m : modes) {
            unsigned long long* out = (m.num == 0) ? c0 : c1;
            k_probe<<<(unsigned)(threads / 128), 128>>>(d, 10, m.num, m.den, sqr, out);
            cudaDeviceSynchronize();
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k_probe<<<(unsigned)(threads / 128), 128>>>(d, iters, m.num, m.den, sqr, out);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            size_t bad = 0;
            if (m.num == 0) cudaMemcpy(href, c0, threads * 8, cudaMemcpyDeviceToHost);
            else { cudaMemcpy(hgot, c1, threads * 8, cudaMemcpyDeviceToHost); for (size_t i = 0; i < threads; i++) bad += href[i] != hgot[i]; }
            printf("%s %-40s %8.3f ms  %7.2f G products/s  mismatches vs integer: %zu  (%s)\n", sqr ? "sqr" : "mul", m.name, ms,
                   threads * 2.0 * iters / (ms * 1e-3) / 1e9, bad, cudaGetErrorString(err));
        }
    }
    return 0;
}
