// lat_probe.cu -- single-warp latency of the Fp product and of a Jacobian doubling in the forms the latency-bound
// kernels could use (tails of the bucket MSM, one-polynomial transforms): dependent chains timed with clock64 on ONE
// warp of one SM, so the numbers are cycles of critical path, not throughput.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I go_kzg_b200/csrc -I include -o tools/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "quad.cuh"
using namespace b200;

__device__ __forceinline__ G1J g1_dbl_inl(const G1J& p) {      // g1_dbl with the products inlined (ILP across them)
    Fp a = fe_sqr(p.x), b = fe_sqr(p.y), yz = fe_mul(p.y, p.z);
    Fp c = fe_sqr(b), t = fe_sqr(fe_add(p.x, b));
    Fp e = fe_add(fe_dbl(a), a);
    Fp f = fe_sqr(e);
    Fp d = fe_dbl(fe_sub(fe_sub(t, a), c));
    G1J r;
    r.z = fe_dbl(yz);
    r.x = fe_sub(f, fe_dbl(d));
    r.y = fe_sub(fe_mul(e, fe_sub(d, r.x)), fe_dbl(fe_dbl(fe_dbl(c))));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(32) k_lat(uint32_t* buf, int iters, long long* cycles, unsigned long long* check) {
    Fp x, y, z;
    for (int k = 0; k < 12; k++) { x.l[k] = buf[k] + threadIdx.x; y.l[k] = buf[12 + k] ^ threadIdx.x; z.l[k] = buf[k] * 3u + threadIdx.x; }
    x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu; z.l[11] &= 0x0fffffffu;
    G1J p; p.x = x; p.y = y; p.z = z;
    long long t0 = clock64();
    if (MODE == 0) for (int k = 0; k < iters; k++) x = fe_mul(x, y);                        // inline, dependent
    if (MODE == 1) for (int k = 0; k < iters; k++) x = fe_sqr(x);
    if (MODE == 2) for (int k = 0; k < iters; k++) x = fp_mul(x, y);                        // out-of-line call
    if (MODE == 3) for (int k = 0; k < iters; k++) { x = fe_mul(x, z); y = fe_mul(y, z); }  // two independent chains
    if (MODE == 4) for (int k = 0; k < iters; k++) { x = fe_sqr(x); y = fe_sqr(y); z = fe_sqr(z); }
    if (MODE == 5) for (int k = 0; k < iters; k++) g1_dbl_ni(&p, &p);                       // what the kernels do today
    if (MODE == 6) for (int k = 0; k < iters; k++) p = g1_dbl_inl(p);
    if (MODE == 7) for (int k = 0; k < iters; k++) x = fe_mul_k(x, y);                      // product then word-serial reduction
    if (MODE == 8) for (int k = 0; k < iters; k++) x = fp_sqr(x);
    if (MODE == 9) for (int k = 0; k < iters; k++) quad_dbl(&p, true);                      // quad operations (quad.cuh)
    if (MODE == 10) { G1J q = p; q.x = y; for (int k = 0; k < iters; k++) quad_add(&p, &q, true); }
    if (MODE == 11) { G1A q; q.x = y; q.y = z; for (int k = 0; k < iters; k++) quad_add_mixed(&p, &q, true); }
    if (MODE == 12) for (int k = 0; k < iters; k++) x = fe_add(x, y);                       // one linear operation, dependent
    if (MODE == 13) for (int k = 0; k < iters; k++) x = quad_bcast(x, (k + threadIdx.x) & 3); // one gather of 12 words
    if (MODE == 14) for (int k = 0; k < iters; k++) x = quad_pick3(quad_role(), x, y, z);
    if (MODE == 15) { G1J q = p; q.x = y; for (int k = 0; k < iters; k++) g1_add_ni(&p, &p, &q); }   // one-lane general addition
    long long t1 = clock64();
    unsigned long long acc = 0;
    for (int k = 0; k < 12; k++) acc = acc * 1000003ull + x.l[k] + 31ull * y.l[k] + 17ull * z.l[k] + p.x.l[k] + p.y.l[k] + p.z.l[k];
    check[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    uint32_t h[24]; for (int i = 0; i < 24; i++) h[i] = 0x9e3779b9u * (i + 1);
    uint32_t* d; cudaMalloc(&d, sizeof(h)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    long long* cyc; cudaMalloc(&cyc, 8);
    unsigned long long* chk; cudaMalloc(&chk, 32 * 8);
    const int iters = 200;
    const char* names[] = {"fe_mul inline, dependent", "fe_sqr inline, dependent", "fp_mul out-of-line call", "2 independent fe_mul per iteration",
                           "3 independent fe_sqr per iteration", "g1_dbl_ni (out-of-line products)", "g1_dbl inlined products",
                           "fe_mul_k (product + serial reduction)", "fp_sqr out-of-line call", "quad_dbl", "quad_add", "quad_add_mixed",
                           "fe_add, dependent", "quad_bcast (12 SHFL)", "quad_pick3", "g1_add_ni (one lane)"};
    for (int mode = 0; mode < 16; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (mode) {
                case 0: k_lat<0><<<1, 32>>>(d, iters, cyc, chk); break;
                case 1: k_lat<1><<<1, 32>>>(d, iters, cyc, chk); break;
                case 2: k_lat<2><<<1, 32>>>(d, iters, cyc, chk); break;
                case 3: k_lat<3><<<1, 32>>>(d, iters, cyc, chk); break;
                case 4: k_lat<4><<<1, 32>>>(d, iters, cyc, chk); break;
                case 5: k_lat<5><<<1, 32>>>(d, iters, cyc, chk); break;
                case 6: k_lat<6><<<1, 32>>>(d, iters, cyc, chk); break;
                case 7: k_lat<7><<<1, 32>>>(d, iters, cyc, chk); break;
                case 8: k_lat<8><<<1, 32>>>(d, iters, cyc, chk); break;
                case 9: k_lat<9><<<1, 32>>>(d, iters, cyc, chk); break;
                case 10: k_lat<10><<<1, 32>>>(d, iters, cyc, chk); break;
                case 11: k_lat<11><<<1, 32>>>(d, iters, cyc, chk); break;
                case 12: k_lat<12><<<1, 32>>>(d, iters, cyc, chk); break;
                case 13: k_lat<13><<<1, 32>>>(d, iters, cyc, chk); break;
                case 14: k_lat<14><<<1, 32>>>(d, iters, cyc, chk); break;
                case 15: k_lat<15><<<1, 32>>>(d, iters, cyc, chk); break;
            }
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-42s %9.1f cycles per iteration  (%s)\n", names[mode], (double)c / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
