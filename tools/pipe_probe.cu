// pipe_probe.cu -- issue-rate microbenchmarks for the integer and FP64 pipes of a B200 SM
// (which multiplier a 381-bit Montgomery product should be built on).  Prints warp-instruction
// issue intervals per SM sub-partition in cycles.   nvcc -arch=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int KIND>
__global__ void __launch_bounds__(256) probe(uint64_t* out, uint32_t seed) {
    uint32_t a[UNROLL], b = seed | 1u;
    uint64_t w[UNROLL];
    double d[UNROLL], e = 1.0000001 + seed * 1e-9, f = 0.25;
    for (int i = 0; i < UNROLL; i++) { a[i] = threadIdx.x * 77u + i + seed; w[i] = a[i]; d[i] = a[i] * 1e-3; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < UNROLL; i++) {
            if (KIND == 0) w[i] = (uint64_t)(uint32_t)w[i] * b + w[i];                         // IMAD.WIDE
            if (KIND == 1) a[i] = a[i] * b + a[i];                                              // IMAD lo
            if (KIND == 2) a[i] = __umulhi(a[i], b) + a[i];                                     // IMAD.HI
            if (KIND == 3) d[i] = __fma_rz(d[i], e, f);                                         // DFMA
            if (KIND == 4) { w[i] = (uint64_t)(uint32_t)w[i] * b + w[i]; d[i] = __fma_rz(d[i], e, f); }   // both pipes
            if (KIND == 5) { d[i] = __fma_rz(d[i], e, f); a[i] = (a[i] + b) ^ (a[i] >> 3); }    // DFMA + 2 ALU ops
            if (KIND == 6) d[i] = d[i] + e;                                                     // DADD
        }
    }
    uint64_t acc = 0;
    for (int i = 0; i < UNROLL; i++) acc += w[i] + a[i] + (uint64_t)__double_as_longlong(d[i]);
    if (acc == 0x1234567ull) out[0] = acc;
}

template <int KIND>
static void run(const char* name, int ops_per_iter, int sms, double mhz) {
    uint64_t* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;   // 8 x 256 threads = 2048 threads per SM (16 warps per sub-partition)
    probe<KIND><<<blocks, 256>>>(d, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<KIND><<<blocks, 256>>>(d, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr_per_smsp = (double)ITERS * UNROLL * ops_per_iter * 16;   // 16 warps per SMSP
    double cycles = ms * 1e-3 * mhz * 1e6;
    printf("%-28s %8.3f ms  %6.2f cycles per warp-instruction per SMSP (at %.0f MHz)\n", name, ms, cycles / warp_instr_per_smsp, mhz);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1e3;
    printf("%s, %d SMs, %.0f MHz nominal\n", p.name, p.multiProcessorCount, mhz);
    run<0>("IMAD.WIDE", 1, p.multiProcessorCount, mhz);
    run<1>("IMAD (lo)", 1, p.multiProcessorCount, mhz);
    run<2>("IMAD.HI", 1, p.multiProcessorCount, mhz);
    run<3>("DFMA", 1, p.multiProcessorCount, mhz);
    run<6>("DADD", 1, p.multiProcessorCount, mhz);
    run<4>("IMAD.WIDE + DFMA (pairs)", 1, p.multiProcessorCount, mhz);
    run<5>("DFMA + 2 ALU (triples)", 1, p.multiProcessorCount, mhz);
    return 0;
}
