// kernels_fr.cu -- scalar-field kernels: radix-2 Cooley-Tukey NTT over Fr (fft_fr.go:30-105),
// Toeplitz coefficient gathers (fk20_single.go:89-119), pointwise helpers.
//
// NTT design: natural order in, natural order out, exactly the reference's transform.  A
// transform of n = 2^logn points runs as one pass (n <= 4096: the whole vector lives in shared
// memory, 32 B per element) or as two passes of a 4-step decomposition n = n1 * n2 (columns of
// length n1 with a twiddle correction, then rows of length n2).  Inside a pass a CTA owns `cols`
// independent sub-transforms: data is gathered with the bit reversal folded into the *global*
// read (32 B elements = whole sectors), the half-table of twiddles w_M^j is staged to shared
// memory with one bulk async copy (cp.async.bulk + mbarrier, TMA's 1-D form), and log2(M)
// butterfly stages run out of shared memory with 128-bit accesses, one butterfly per lane.
#include "field.cuh"
#include "kernels.h"

namespace b200 {

// Function attributes belong to a device's context: remember per (kernel, device) whether the large dynamic
// shared-memory opt-in has been made (a process may drive several GPUs through b200_set_device).
static bool smem_attr_done(int which) {
    static bool done[2][64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    bool was = done[which][dev];
    done[which][dev] = true;
    return was;
}


static inline unsigned grid_for(size_t total, unsigned block) { return (unsigned)((total + block - 1) / block); }
__device__ __forceinline__ uint32_t brev_bits(uint32_t v, unsigned logn) { return logn ? (__brev(v) >> (32 - logn)) : 0u; }

// ------------------------------------------------------------------------------ conversions
__global__ void k_fr_to_mont(const uint64_t* __restrict__ in, Fr* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr a = ld_vec(reinterpret_cast<const Fr*>(in) + i);
    st_vec(out + i, fe_to_mont(a));
}
__global__ void k_fr_from_mont(const Fr* __restrict__ in, uint64_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_vec(reinterpret_cast<Fr*>(out) + i, fe_from_mont(ld_vec(in + i)));
}
void launch_fr_to_mont(const uint64_t* in, Fr* out, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_fr_to_mont<<<grid_for(n, 256), 256, 0, st>>>(in, out, n); g_launch_count++;
}
void launch_fr_from_mont(const Fr* in, uint64_t* out, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_fr_from_mont<<<grid_for(n, 256), 256, 0, st>>>(in, out, n); g_launch_count++;
}

__global__ void k_fr_mul_arrays(Fr* dst, const Fr* a, const Fr* b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_vec(dst + i, fe_mul(ld_vec(a + i), ld_vec(b + i)));
}
void launch_fr_mul_arrays(Fr* dst, const Fr* a, const Fr* b, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_fr_mul_arrays<<<grid_for(n, 256), 256, 0, st>>>(dst, a, b, n); g_launch_count++;
}

// ------------------------------------------------------------------------------ NTT pass
struct NttPass {
    const Fr* in;
    Fr* out;
    size_t n;             // points per transform
    unsigned logm;        // sub-transform length M = 2^logm
    unsigned cols;        // sub-transforms per CTA (power of two)
    size_t in_estride, in_cstride;     // element e of column c: in[b n + e in_estride + c in_cstride]
    size_t out_kstride, out_cstride;   // output k of column c: out[b n + k out_kstride + c out_cstride]
    const Fr* tw;         // M/2 twiddles w_M^j (forward or inverse), contiguous
    const Fr* big;        // w_N^i table (expanded or reverse roots), or null: no correction
    size_t big_stride;    // max_width / n
    int c_fast;           // gather order of the global read (1: column index fastest)
    int has_scale;
    Fr scale;
};

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    uint32_t ok = 0;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(phase)
            : "memory");
    }
}

__global__ void __launch_bounds__(1024) k_fr_ntt_pass(NttPass P) {
    extern __shared__ uint4 smem_raw[];
    __shared__ uint64_t bar;
    const unsigned M = 1u << P.logm, C = P.cols;
    Fr* s = reinterpret_cast<Fr*>(smem_raw);    // [M][C]
    Fr* tws = s + (size_t)M * C;                // [M/2]
    const unsigned tid = threadIdx.x, T = blockDim.x;
    const size_t c0 = (size_t)blockIdx.x * C;
    const Fr* in = P.in + (size_t)blockIdx.y * P.n;
    Fr* out = P.out + (size_t)blockIdx.y * P.n;

    const uint32_t tw_bytes = (M / 2) * (uint32_t)sizeof(Fr);
    if (tid == 0 && tw_bytes) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar, tw_bytes);
        bulk_g2s(tws, P.tw, tw_bytes, &bar);
    }
    // gather with the bit reversal on the global side: s[p][c] = x[rev(p)][c]
    const unsigned total = M * C;
    for (unsigned idx = tid; idx < total; idx += T) {
        unsigned p, c;
        if (P.c_fast) { c = idx % C; p = idx / C; } else { p = idx % M; c = idx / M; }
        unsigned e = brev_bits(p, P.logm);
        st_vec(s + (size_t)p * C + c, ld_vec(in + (size_t)e * P.in_estride + (c0 + c) * P.in_cstride));
    }
    __syncthreads();
    if (tw_bytes) mbar_wait(&bar, 0);   // every thread observes the completed phase (acquire)

    // Stages in pairs (radix 4): a thread takes the four elements g, g + m, g + 2m, g + 3m through the stages
    // m and 2m in registers -- the same four twiddle products as two radix-2 stages, half the shared-memory
    // traffic and half the barriers.  An odd stage count starts with one radix-2 stage.
    unsigned m = 1;
    if (P.logm & 1) {
        const unsigned nbf = (M / 2) * C;
        for (unsigned bf = tid; bf < nbf; bf += T) {          // m = 1: j = 0, no twiddle
            unsigned c = bf % C, q = bf / C;
            Fr* a0 = s + (size_t)(2 * q) * C + c;
            Fr* a1 = a0 + C;
            Fr x0 = ld_vec(a0), x1 = ld_vec(a1);
            st_vec(a0, fe_add(x0, x1));
            st_vec(a1, fe_sub(x0, x1));
        }
        __syncthreads();
        m = 2;
    }
    const unsigned nunits = (M / 4) * C;
    for (; m < M; m <<= 2) {
        const unsigned ts1 = (M / 2) / m, ts2 = ts1 / 2;
        for (unsigned u = tid; u < nunits; u += T) {
            unsigned c = u % C, q = u / C;
            unsigned j = q & (m - 1);
            unsigned g = ((q - j) << 2) + j;
            Fr* a0 = s + (size_t)g * C + c;
            Fr* a1 = a0 + (size_t)m * C;
            Fr* a2 = a1 + (size_t)m * C;
            Fr* a3 = a2 + (size_t)m * C;
            Fr x0 = ld_vec(a0), x1 = ld_vec(a1), x2 = ld_vec(a2), x3 = ld_vec(a3);
            if (j) {
                Fr w = ld_vec(tws + (size_t)j * ts1);
                x1 = fe_mul(x1, w);
                x3 = fe_mul(x3, w);
            }
            Fr y0 = fe_add(x0, x1), y1 = fe_sub(x0, x1), y2 = fe_add(x2, x3), y3 = fe_sub(x2, x3);
            if (j) y2 = fe_mul(y2, ld_vec(tws + (size_t)j * ts2));
            y3 = fe_mul(y3, ld_vec(tws + (size_t)(j + m) * ts2));
            st_vec(a0, fe_add(y0, y2));
            st_vec(a2, fe_sub(y0, y2));
            st_vec(a1, fe_add(y1, y3));
            st_vec(a3, fe_sub(y1, y3));
        }
        __syncthreads();
    }
    for (unsigned idx = tid; idx < total; idx += T) {
        unsigned c = idx % C, k = idx / C;
        Fr v = ld_vec(s + (size_t)k * C + c);
        if (P.big) {
            size_t ti = (size_t)k * (c0 + c);
            if (ti) v = fe_mul(v, ld_vec(P.big + ti * P.big_stride));
        }
        if (P.has_scale) v = fe_mul(v, P.scale);
        st_vec(out + (size_t)k * P.out_kstride + (c0 + c) * P.out_cstride, v);
    }
}

static void run_pass(NttPass& P, size_t ngroups, size_t batch, cudaStream_t st) {
    const size_t M = (size_t)1 << P.logm;
    size_t smem = M * P.cols * sizeof(Fr) + (M / 2) * sizeof(Fr);
    size_t work = M * P.cols / 4;       // radix-4 units per stage pair
    unsigned threads = work >= 1024 ? 1024 : (work < 32 ? 32 : (unsigned)work);   // 64 registers: measured faster than 512 x 92
    if (!smem_attr_done(0)) cudaFuncSetAttribute(k_fr_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    // the batch rides on grid.y (at most 65535): larger batches (the product tree of the zero polynomial has
    // batch * roots / 32 nodes) go in chunks
    const Fr* in0 = P.in;
    Fr* out0 = P.out;
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {
        const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        P.in = in0 + b0 * P.n; P.out = out0 + b0 * P.n;
        dim3 grid((unsigned)ngroups, (unsigned)nb);
        k_fr_ntt_pass<<<grid, threads, smem, st>>>(P);
        g_launch_count++;
    }
    P.in = in0; P.out = out0;
}

void launch_fr_ntt(const FrDomain& dom, const Fr* in, Fr* out, Fr* tmp, unsigned logn, size_t batch, bool inverse,
                   const Fr* scale_or_null, cudaStream_t st) {
    ProfScope prof_scope(PROF_FR_NTT, st);
    if (!batch) return;
    const size_t n = (size_t)1 << logn;
    const Fr* twbase = inverse ? dom.tw_inv : dom.tw_fwd;
    const Fr* big = inverse ? dom.reverse : dom.expanded;
    NttPass P;
    P.n = n;
    P.has_scale = 0;
    if (logn <= 12) {
        P.logm = logn;
        P.in_estride = 1; P.out_kstride = 1;
        P.tw = twbase + n / 2; P.big = nullptr; P.big_stride = 0; P.c_fast = 0;
        if (scale_or_null) { P.has_scale = 1; P.scale = *scale_or_null; }
        // Many short transforms (the levels of the zero-polynomial product tree, interpolations of CheckProofMulti): one CTA per
        // transform would be a handful of threads each; pack G transforms into the columns of one CTA's tile instead.
        size_t done = 0;
        if (logn <= 10 && batch >= 4) {
            size_t G = 4096 / n;
            if (G > 64) G = 64;
            while (G > batch) G >>= 1;
            const size_t groups = batch / G;
            P.in = in; P.out = out; P.cols = (unsigned)G;
            P.in_cstride = n; P.out_cstride = n;          // column c of group g = transform g G + c
            for (size_t g0 = 0; g0 < groups; g0 += 65535) {   // grid.x carries the groups, grid.y = 1
                const size_t ng = groups - g0 < 65535 ? groups - g0 : 65535;
                NttPass Q = P;
                Q.in = in + g0 * G * n; Q.out = out + g0 * G * n;
                run_pass(Q, ng, 1, st);
            }
            done = groups * G;
        }
        if (done < batch) {
            P.in = in + done * n; P.out = out + done * n; P.cols = 1;
            P.in_cstride = 0; P.out_cstride = 0;
            run_pass(P, 1, batch - done, st);
        }
        return;
    }
    const unsigned log1 = (logn + 1) / 2, log2 = logn - log1;
    const size_t n1 = (size_t)1 << log1, n2 = (size_t)1 << log2;
    // pass 1: columns (length n1, stride n2), times w_n^(n2 k1), left in place at [k1][n2]
    {
        unsigned cols = (unsigned)(4096 / n1); if (cols > 8) cols = 8; if (cols > n2) cols = (unsigned)n2;
        P.in = in; P.out = tmp; P.logm = log1; P.cols = cols;
        P.in_estride = n2; P.in_cstride = 1; P.out_kstride = n2; P.out_cstride = 1;
        P.tw = twbase + n1 / 2; P.big = big; P.big_stride = dom.max_width / n; P.c_fast = 1;
        run_pass(P, n2 / cols, batch, st);
    }
    // pass 2: rows (length n2, contiguous), output k2 of row k1 lands at k1 + n1 k2
    {
        unsigned cols = (unsigned)(4096 / n2); if (cols > 8) cols = 8; if (cols > n1) cols = (unsigned)n1;
        P.in = tmp; P.out = out; P.logm = log2; P.cols = cols;
        P.in_estride = 1; P.in_cstride = n2; P.out_kstride = n1; P.out_cstride = 1;
        P.tw = twbase + n2 / 2; P.big = nullptr; P.big_stride = 0; P.c_fast = 0;
        if (scale_or_null) { P.has_scale = 1; P.scale = *scale_or_null; }
        run_pass(P, n1 / cols, batch, st);
    }
}

// ------------------------------------------------------------------------------ DAS extension
// das_extension.go:7-66 as an in-place butterfly network over n = 2^logn values.  Root indices
// are relative to the settings' full domain exactly as in the reference (domainStride starts at
// 1 whatever n is).
//   descent, block length L = n .. 4 (stride s = n / L):  (a0, a1) <- (a0 + a1, (a0 - a1) Rev[2 i s])
//   base,    L = 2 (s = n / 2):                            x = a0 + a1, t = (a0 - a1) Exp[s]; (x + t, x - t)
//   ascent,  L = 4 .. n:                                   (a0, a1) <- (a0 + a1 Exp[(1 + 2 i) s], a0 - a1 Exp[..])
//   finally every value times n^-1                         das_extension.go:78-83
// Blocks of up to DAS_BLOCK values run entirely in shared memory; longer levels are single
// global-memory passes.
#define DAS_LOG_BLOCK 12
template <bool ASCENT>
__global__ void k_das_level(Fr* vals, size_t n, size_t L, const Fr* __restrict__ roots, int has_scale, Fr scale) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fr* v = vals + (size_t)blockIdx.y * n;
    if (t >= n / 2) return;
    size_t hh = L / 2, i = t % hh, base = (t / hh) * L, s = n / L;
    Fr a0 = ld_vec(v + base + i), a1 = ld_vec(v + base + hh + i);
    if (!ASCENT) {
        Fr d = fe_mul(fe_sub(a0, a1), ld_vec(roots + 2 * i * s));
        st_vec(v + base + i, fe_add(a0, a1));
        st_vec(v + base + hh + i, d);
    } else {
        Fr yr = fe_mul(a1, ld_vec(roots + (1 + 2 * i) * s));
        Fr r0 = fe_add(a0, yr), r1 = fe_sub(a0, yr);
        if (has_scale) { r0 = fe_mul(r0, scale); r1 = fe_mul(r1, scale); }
        st_vec(v + base + i, r0);
        st_vec(v + base + hh + i, r1);
    }
}
// one CTA = one block of B = 2^logb consecutive values (all levels L <= B)
__global__ void __launch_bounds__(1024) k_das_block(Fr* vals, size_t n, unsigned logb, const Fr* __restrict__ expanded,
                                                    const Fr* __restrict__ reverse, int has_scale, Fr scale) {
    extern __shared__ uint4 smem_raw[];
    Fr* s = reinterpret_cast<Fr*>(smem_raw);
    const unsigned B = 1u << logb, tid = threadIdx.x, T = blockDim.x;
    Fr* v = vals + (size_t)blockIdx.y * n + (size_t)blockIdx.x * B;
    for (unsigned i = tid; i < B; i += T) st_vec(s + i, ld_vec(v + i));
    __syncthreads();
    for (unsigned L = B; L >= 4; L >>= 1) {
        const unsigned hh = L / 2;
        const size_t st = n / L;
        for (unsigned t = tid; t < B / 2; t += T) {
            unsigned i = t % hh, base = (t / hh) * L;
            Fr a0 = ld_vec(s + base + i), a1 = ld_vec(s + base + hh + i);
            Fr d = fe_mul(fe_sub(a0, a1), ld_vec(reverse + 2 * (size_t)i * st));
            st_vec(s + base + i, fe_add(a0, a1));
            st_vec(s + base + hh + i, d);
        }
        __syncthreads();
    }
    if (B >= 2) {
        const Fr w = ld_vec(expanded + n / 2);
        for (unsigned t = tid; t < B / 2; t += T) {
            Fr a0 = ld_vec(s + 2 * t), a1 = ld_vec(s + 2 * t + 1);
            Fr x = fe_add(a0, a1), y = fe_mul(fe_sub(a0, a1), w);
            st_vec(s + 2 * t, fe_add(x, y));
            st_vec(s + 2 * t + 1, fe_sub(x, y));
        }
        __syncthreads();
    }
    for (unsigned L = 4; L <= B; L <<= 1) {
        const unsigned hh = L / 2;
        const size_t st = n / L;
        for (unsigned t = tid; t < B / 2; t += T) {
            unsigned i = t % hh, base = (t / hh) * L;
            Fr a0 = ld_vec(s + base + i), a1 = ld_vec(s + base + hh + i);
            Fr yr = fe_mul(a1, ld_vec(expanded + (1 + 2 * (size_t)i) * st));
            st_vec(s + base + i, fe_add(a0, yr));
            st_vec(s + base + hh + i, fe_sub(a0, yr));
        }
        __syncthreads();
    }
    for (unsigned i = tid; i < B; i += T) {
        Fr x = ld_vec(s + i);
        if (has_scale) x = fe_mul(x, scale);
        st_vec(v + i, x);
    }
}
void launch_das_fft_extension(const FrDomain& dom, Fr* vals, unsigned logn, size_t batch, const Fr& inv_n, cudaStream_t st) {
    ProfScope prof_scope(PROF_FR_NTT, st);
    if (!batch) return;
    const size_t n = (size_t)1 << logn;
    if (logn == 0) return;   // the reference panics ("bad usage") before this point; caller checks
    const unsigned logb = logn < DAS_LOG_BLOCK ? logn : DAS_LOG_BLOCK;
    const size_t B = (size_t)1 << logb;
    if (!smem_attr_done(1)) cudaFuncSetAttribute(k_das_block, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 lgrid((unsigned)((n / 2 + 255) / 256), (unsigned)batch);
    for (size_t L = n; L > B; L >>= 1) { k_das_level<false><<<lgrid, 256, 0, st>>>(vals, n, L, dom.reverse, 0, inv_n); g_launch_count++; }
    unsigned threads = B / 2 >= 1024 ? 1024 : (B / 2 < 32 ? 32 : (unsigned)(B / 2));
    k_das_block<<<dim3((unsigned)(n / B), (unsigned)batch), threads, B * sizeof(Fr), st>>>(vals, n, logb, dom.expanded, dom.reverse,
                                                                                    B == n ? 1 : 0, inv_n);
    g_launch_count++;
    for (size_t L = 2 * B; L <= n; L <<= 1) { k_das_level<true><<<lgrid, 256, 0, st>>>(vals, n, L, dom.expanded, L == n ? 1 : 0, inv_n); g_launch_count++; }
}

// ------------------------------------------------------------------------------ zero polynomial
// zero_poly.go:116-217 computes Z(X) = prod_i (X - w^{m_i}) over the missing indices with a leaf /
// FFT-convolution tree and then evaluates it.  The result is a uniquely determined polynomial,
// so the device takes the shortest exact route instead: evaluate the product directly,
//      zeroEval[j] = prod_i (w^j - w^{m_i}),
// every (j, root-segment) pair in parallel with the segment's roots staged in shared memory, and
// recover the coefficients with one inverse NTT (deg Z < length whenever the reference does not
// panic).  Same field elements, hence the same bytes.
#define ZP_SEG 512
__global__ void __launch_bounds__(256) k_zero_eval_partial(const Fr* __restrict__ expanded, size_t stride, size_t n,
                                                           const uint32_t* __restrict__ missing, const uint32_t* __restrict__ nmiss,
                                                           size_t miss_pitch, Fr* __restrict__ partial, size_t nseg) {
    __shared__ uint4 roots_raw[ZP_SEG * 2];
    Fr* roots = reinterpret_cast<Fr*>(roots_raw);
    const size_t b = blockIdx.z, seg = blockIdx.y;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t cnt_all = nmiss[b];
    const size_t lo = seg * ZP_SEG;
    const uint32_t cnt = lo >= cnt_all ? 0u : (cnt_all - lo > ZP_SEG ? (uint32_t)ZP_SEG : (uint32_t)(cnt_all - lo));
    for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x)
        st_vec(roots + i, ld_vec(expanded + (size_t)missing[b * miss_pitch + lo + i] * stride));
    __syncthreads();
    if (j >= n) return;
    const Fr x = ld_vec(expanded + j * stride);
    Fr acc = Fr::one();
    for (unsigned i = 0; i < cnt; i++) acc = fe_mul(acc, fe_sub(x, ld_vec(roots + i)));
    st_vec(partial + (b * nseg + seg) * n + j, acc);
}
__global__ void k_zero_eval_combine(const Fr* __restrict__ partial, Fr* __restrict__ zero_eval, size_t n, size_t nseg,
                                    const uint32_t* __restrict__ nmiss) {
    const size_t b = blockIdx.y;
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (nmiss[b] == 0) { st_vec(zero_eval + b * n + j, Fr::zero()); return; }   // zero_poly.go:117-119
    Fr acc = ld_vec(partial + (b * nseg) * n + j);
    for (size_t s = 1; s < nseg; s++) acc = fe_mul(acc, ld_vec(partial + (b * nseg + s) * n + j));
    st_vec(zero_eval + b * n + j, acc);
}
void launch_zero_eval(const FrDomain& dom, size_t n, size_t batch, const uint32_t* d_missing, const uint32_t* d_nmiss,
                      size_t miss_pitch, size_t max_missing, Fr* partial, Fr* zero_eval, cudaStream_t st) {
    ProfScope prof_scope(PROF_FR_NTT, st);
    if (!n || !batch) return;
    size_t nseg = (max_missing + ZP_SEG - 1) / ZP_SEG;
    if (nseg == 0) nseg = 1;
    dim3 g1((unsigned)((n + 255) / 256), (unsigned)nseg, (unsigned)batch);
    k_zero_eval_partial<<<g1, 256, 0, st>>>(dom.expanded, dom.max_width / n, n, d_missing, d_nmiss, miss_pitch, partial, nseg);
    dim3 g2((unsigned)((n + 255) / 256), (unsigned)batch);
    k_zero_eval_combine<<<g2, 256, 0, st>>>(partial, zero_eval, n, nseg, d_nmiss);
    g_launch_count += 2;
}
size_t zero_eval_segments(size_t max_missing) { size_t s = (max_missing + ZP_SEG - 1) / ZP_SEG; return s ? s : 1; }

// ---- product tree for large missing sets --------------------------------------------------------------
// The direct evaluation above costs n * |missing| products (1.3 * 10^8 at n = 2^14, half missing).  For large
// sets the coefficients of Z are built like zero_poly.go:17-113 does -- leaves of ZP_LEAF roots multiplied out
// directly, then pairwise products through NTTs -- and evaluated with one forward NTT.  Differences that do not
// change the polynomial: binary instead of 4-way reduction, monic factors stored without their leading 1
// ((X^d + p)(X^d + q) = X^2d + X^d (p + q) + p q, so a size-2d cyclic convolution is exact), and every
// list padded to ZP_LEAF * 2^L roots with the root 0 (a factor X^k that is shifted out at the end).
#define ZP_LEAF 32
// coef[b][leaf][0..32): low coefficients of prod_i (X - r_i), r_i = w^missing or 0 (padding)
__global__ void __launch_bounds__(128) k_zp_leaves(const Fr* __restrict__ expanded, size_t stride, const uint32_t* __restrict__ missing,
                                                   const uint32_t* __restrict__ nmiss, size_t miss_pitch, size_t leaves, size_t batch,
                                                   Fr* __restrict__ coef) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= leaves * batch) return;
    size_t b = t / leaves, leaf = t % leaves;
    const uint32_t cnt = nmiss[b];
    Fr c[ZP_LEAF];
    for (int deg = 0; deg < ZP_LEAF; deg++) {
        size_t ri = leaf * ZP_LEAF + deg;
        Fr r = ri < cnt ? ld_vec(expanded + (size_t)missing[b * miss_pitch + ri] * stride) : Fr::zero();
        // (X^deg + c) (X - r): new c_deg = c_(deg-1) - r, c_k = c_(k-1) - r c_k, c_0 = -r c_0
        if (deg == 0) { c[0] = fe_neg(r); continue; }
        c[deg] = fe_sub(c[deg - 1], r);
        for (int k = deg - 1; k >= 1; k--) c[k] = fe_sub(c[k - 1], fe_mul(r, c[k]));
        c[0] = fe_neg(fe_mul(r, c[0]));
    }
    for (int k = 0; k < ZP_LEAF; k++) st_vec(coef + t * ZP_LEAF + k, c[k]);
}
// padded[node][0..d) = coef[node][0..d), padded[node][d..2d) = 0
__global__ void k_zp_pad(const Fr* __restrict__ coef, Fr* __restrict__ padded, size_t d, size_t total /* nodes * 2d */) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    size_t node = t / (2 * d), k = t % (2 * d);
    st_vec(padded + t, k < d ? ld_vec(coef + node * d + k) : Fr::zero());
}
// prod[pair][k] = padded[2 pair][k] * padded[2 pair + 1][k]   (transform domain, 2d values per node)
__global__ void k_zp_pointwise(const Fr* __restrict__ padded, Fr* __restrict__ prod, size_t d2, size_t total /* pairs * 2d */) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    size_t pair = t / d2, k = t % d2;
    st_vec(prod + t, fe_mul(ld_vec(padded + (2 * pair) * d2 + k), ld_vec(padded + (2 * pair + 1) * d2 + k)));
}
// next[pair][k] = (p q)[k] + (k >= d ? p[k - d] + q[k - d] : 0)
__global__ void k_zp_combine(const Fr* __restrict__ prod, const Fr* __restrict__ coef, Fr* __restrict__ next, size_t d,
                             size_t total /* pairs * 2d */) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    size_t pair = t / (2 * d), k = t % (2 * d);
    Fr v = ld_vec(prod + t);
    if (k >= d) v = fe_add(v, fe_add(ld_vec(coef + (2 * pair) * d + (k - d)), ld_vec(coef + (2 * pair + 1) * d + (k - d))));
    st_vec(next + t, v);
}
// zero_poly[b][i] = coefficient i + (mp - nmiss[b]) of X^mp + coef[b]   (i < n); all zero if nothing is missing
__global__ void k_zp_finish(const Fr* __restrict__ coef, const uint32_t* __restrict__ nmiss, size_t mp, size_t n, size_t batch,
                            Fr* __restrict__ zero_poly) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    size_t b = t / n, i = t % n;
    const uint32_t cnt = nmiss[b];
    size_t src = i + (mp - cnt);
    Fr v = Fr::zero();
    if (cnt) { if (src < mp) v = ld_vec(coef + b * mp + src); else if (src == mp) v = Fr::one(); }
    st_vec(zero_poly + t, v);
}
size_t zero_poly_tree_size(size_t max_missing) {       // padded root count ZP_LEAF * 2^L
    size_t mp = ZP_LEAF;
    while (mp < max_missing) mp <<= 1;
    return mp;
}
void launch_zero_poly_tree(const FrDomain& dom, size_t n, size_t batch, const uint32_t* d_missing, const uint32_t* d_nmiss,
                           size_t miss_pitch, size_t mp, Fr* coef_a /* batch mp */, Fr* coef_b /* batch mp */,
                           Fr* padded /* batch 2 mp */, Fr* ntt_tmp /* batch 2 mp */, Fr* zero_poly /* batch n */, cudaStream_t st) {
    {
        ProfScope prof_scope(PROF_MISC, st);
        size_t leaves = mp / ZP_LEAF;
        k_zp_leaves<<<grid_for(leaves * batch, 128), 128, 0, st>>>(dom.expanded, dom.max_width / n, d_missing, d_nmiss, miss_pitch, leaves, batch, coef_a);
        g_launch_count++;
    }
    Fr* cur = coef_a; Fr* nxt = coef_b;
    for (size_t d = ZP_LEAF; d < mp; d <<= 1) {
        const size_t nodes = batch * (mp / d), pairs = nodes / 2, d2 = 2 * d;
        unsigned logd2 = 0; while (((size_t)1 << logd2) < d2) logd2++;
        {
            ProfScope prof_scope(PROF_MISC, st);
            k_zp_pad<<<grid_for(nodes * d2, 256), 256, 0, st>>>(cur, padded, d, nodes * d2); g_launch_count++;
        }
        launch_fr_ntt(dom, padded, padded, ntt_tmp, logd2, nodes, false, nullptr, st);
        {
            ProfScope prof_scope(PROF_MISC, st);
            k_zp_pointwise<<<grid_for(pairs * d2, 256), 256, 0, st>>>(padded, nxt, d2, pairs * d2); g_launch_count++;
        }
        Fr inv = fe_inv(fe_to_mont([&] { Fr v = Fr::zero(); v.l[0] = (uint32_t)d2; v.l[1] = (uint32_t)((uint64_t)d2 >> 32); return v; }()));
        launch_fr_ntt(dom, nxt, nxt, ntt_tmp, logd2, pairs, true, &inv, st);
        {
            ProfScope prof_scope(PROF_MISC, st);
            k_zp_combine<<<grid_for(pairs * d2, 256), 256, 0, st>>>(nxt, cur, padded, d, pairs * d2); g_launch_count++;
        }
        // the combined level sits in `padded` (first batch * mp elements): make it the current compact buffer
        cudaMemcpyAsync(cur, padded, pairs * d2 * sizeof(Fr), cudaMemcpyDeviceToDevice, st);
    }
    ProfScope prof_scope(PROF_MISC, st);
    k_zp_finish<<<grid_for(n * batch, 256), 256, 0, st>>>(cur, d_nmiss, mp, n, batch, zero_poly); g_launch_count++;
    (void)nxt;
}

// ------------------------------------------------------------------------------ recovery helpers
// dst[b][i] = present[b][i] ? a[b][i] * c[b][i] : 0          recover_from_samples.go:60-67
__global__ void k_fr_mul_masked(Fr* dst, const Fr* a, const Fr* c, const uint8_t* __restrict__ present, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Fr v = Fr::zero();
    if (present[i]) v = fe_mul(ld_vec(a + i), ld_vec(c + i));
    st_vec(dst + i, v);
}
void launch_fr_mul_masked(Fr* dst, const Fr* a, const Fr* c, const uint8_t* present, size_t total, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!total) return;
    k_fr_mul_masked<<<grid_for(total, 256), 256, 0, st>>>(dst, a, c, present, total); g_launch_count++;
}
// v[b][i] *= table[i]     (ShiftPoly / UnshiftPoly with precomputed powers, recover_from_samples.go:9-40)
__global__ void k_fr_mul_table(Fr* v, const Fr* __restrict__ table, size_t n, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    st_vec(v + i, fe_mul(ld_vec(v + i), ld_vec(table + (i % n))));
}
// v[i] *= (i even ? even : odd): constants by value, nothing to stage on the device
__global__ void k_fr_mul_even_odd(Fr* v, Fr even, Fr odd, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    st_vec(v + i, fe_mul(ld_vec(v + i), (i & 1) ? odd : even));
}
void launch_fr_mul_even_odd(Fr* v, const Fr& even, const Fr& odd, size_t total, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!total) return;
    k_fr_mul_even_odd<<<grid_for(total, 256), 256, 0, st>>>(v, even, odd, total); g_launch_count++;
}
void launch_fr_mul_table(Fr* v, const Fr* table, size_t n, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_fr_mul_table<<<grid_for(n * batch, 256), 256, 0, st>>>(v, table, n, n * batch); g_launch_count++;
}
// a[i] = a[i] / c[i] with Montgomery's trick over DIV_CHUNK consecutive elements per lane: one
// Fermat inversion per chunk instead of one per element (the reference inverts every element,
// recover_from_samples.go:89-91 / bls/bignum_kilic.go:103-107; x / 0 = 0 there as here).
// CHUNK elements per lane: the Fermat inversion (~380 products) is shared by CHUNK divisions (3 products each).  16 keeps many
// lanes busy for a single polynomial; 64 quarters the inversions for batches (recovery n = 2^14 x 64: 0.71 -> measured below).
template <int CHUNK>
__global__ void __launch_bounds__(128) k_fr_div(Fr* a, const Fr* __restrict__ c, size_t total) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = t * CHUNK;
    if (lo >= total) return;
    size_t cnt = total - lo < CHUNK ? total - lo : CHUNK;
    Fr pre[CHUNK];
    Fr acc = Fr::one();
    for (size_t i = 0; i < cnt; i++) {
        pre[i] = acc;
        Fr d = ld_vec(c + lo + i);
        if (!d.is_zero()) acc = fe_mul(acc, d);
    }
    acc = fe_inv(acc);
    for (size_t i = cnt; i-- > 0;) {
        Fr d = ld_vec(c + lo + i);
        if (d.is_zero()) { st_vec(a + lo + i, Fr::zero()); continue; }
        Fr inv = fe_mul(acc, pre[i]);
        acc = fe_mul(acc, d);
        st_vec(a + lo + i, fe_mul(ld_vec(a + lo + i), inv));
    }
}
void launch_fr_div(Fr* a, const Fr* c, size_t total, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!total) return;
    if (total >= ((size_t)1 << 18)) k_fr_div<64><<<grid_for((total + 63) / 64, 128), 128, 0, st>>>(a, c, total);
    else k_fr_div<16><<<grid_for((total + 15) / 16, 128), 128, 0, st>>>(a, c, total);
    g_launch_count++;
}
// flags[b] |= 1 if a known sample changed (recover_from_samples.go:103-107);
// flags[b] |= 2 if zeroEval[i] == 0 disagrees with "missing" (recover_from_samples.go:54-58)
__global__ void k_recover_check(const Fr* __restrict__ rec, const Fr* __restrict__ samples, const Fr* __restrict__ zero_eval,
                                const uint8_t* __restrict__ present, size_t n, size_t total, uint32_t* flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    unsigned f = 0;
    bool pr = present[i] != 0;
    if (pr && ld_vec(rec + i) != ld_vec(samples + i)) f |= 1;
    if ((!pr) != ld_vec(zero_eval + i).is_zero()) f |= 2;
    if (f) atomicOr(flags + i / n, f);
}
void launch_recover_check(const Fr* rec, const Fr* samples, const Fr* zero_eval, const uint8_t* present, size_t n, size_t batch,
                          uint32_t* flags, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_recover_check<<<grid_for(n * batch, 256), 256, 0, st>>>(rec, samples, zero_eval, present, n, n * batch, flags); g_launch_count++;
}

// ------------------------------------------------------------------------------ missing index lists
// recover_from_samples.go:45-50 on the device: for every polynomial the ascending list of indices i with present[i] == 0
// (stream compaction: one CTA per polynomial, every thread owns n / blockDim consecutive entries; block-wide exclusive scan
// of the per-thread counts).  On the host this loop cost more than the whole device pipeline of a batch (config 4).
__global__ void __launch_bounds__(1024) k_missing_lists(const uint8_t* __restrict__ present, size_t n, uint32_t* __restrict__ missing,
                                                        size_t pitch, uint32_t* __restrict__ nmiss) {
    __shared__ uint32_t warp_tot[32];
    const size_t b = blockIdx.x;
    const uint8_t* pr = present + b * n;
    const unsigned nt = blockDim.x, tid = threadIdx.x;
    const size_t per = (n + nt - 1) / nt, lo = (size_t)tid * per, hi = lo + per < n ? lo + per : n;
    uint32_t c = 0;
    for (size_t i = lo; i < hi; i++) c += pr[i] == 0;
    // inclusive scan inside the warp, then over the warp totals
    uint32_t incl = c;
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, d); if ((tid & 31) >= (unsigned)d) incl += v; }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        uint32_t t = tid < (nt + 31) / 32 ? warp_tot[tid] : 0u, ti = t;
        for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, ti, d); if (tid >= (unsigned)d) ti += v; }
        warp_tot[tid] = ti - t;                                   // exclusive prefix of the warp totals
        if (tid == 31) nmiss[b] = ti;
    }
    __syncthreads();
    uint32_t pos = warp_tot[tid >> 5] + incl - c;
    uint32_t* out = missing + b * pitch;
    for (size_t i = lo; i < hi; i++) if (pr[i] == 0) out[pos++] = (uint32_t)i;
}
void launch_missing_lists(const uint8_t* present, size_t n, size_t batch, uint32_t* missing, size_t pitch, uint32_t* nmiss, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_missing_lists<<<(unsigned)batch, 1024, 0, st>>>(present, n, missing, pitch, nmiss); g_launch_count++;
}

// ------------------------------------------------------------------------------ powers
// out[i] = s^i (canonical), i < n, by square-and-multiply over the bits of i; sq[j] = s^(2^j) (Montgomery)
__global__ void k_fr_powers(const Fr* __restrict__ sq, Fr* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr acc = Fr::one();
    for (int j = 0; j < 40; j++) if ((i >> j) & 1) acc = fe_mul(acc, ld_vec(sq + j));
    st_vec(out + i, fe_from_mont(acc));
}
void launch_fr_powers(const Fr* d_sq, Fr* out_canon, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_fr_powers<<<grid_for(n, 256), 256, 0, st>>>(d_sq, out_canon, n); g_launch_count++;
}

// ------------------------------------------------------------------------------ Toeplitz gathers
// fk20_single.go:106-119: [p[n-1], 0 x (n+1), p[1..n-2]]
__global__ void k_toeplitz_coeffs(const uint64_t* __restrict__ polys, Fr* __restrict__ out, size_t n, size_t batch) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n2 = 2 * n;
    if (t >= n2 * batch) return;
    size_t b = t / n2, i = t % n2;
    const Fr* p = reinterpret_cast<const Fr*>(polys) + b * n;
    Fr v = Fr::zero();
    if (i == 0) v = fe_to_mont(ld_vec(p + (n - 1)));
    else if (i >= n + 2) v = fe_to_mont(ld_vec(p + (i - n - 1)));
    st_vec(out + t, v);
}
void launch_toeplitz_coeffs(const uint64_t* polys, Fr* out, size_t n, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_toeplitz_coeffs<<<grid_for(2 * n * batch, 256), 256, 0, st>>>(polys, out, n, batch); g_launch_count++;
}
// fk20_single.go:89-103 for every offset: out[b][off][2k]
__global__ void k_toeplitz_coeffs_strided(const uint64_t* __restrict__ polys, Fr* __restrict__ out, size_t n, size_t l,
                                          size_t batch) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t k = n / l, k2 = 2 * k;
    if (t >= k2 * l * batch) return;
    size_t i = t % k2, off = (t / k2) % l, b = t / (k2 * l);
    const Fr* p = reinterpret_cast<const Fr*>(polys) + b * n;
    Fr v = Fr::zero();
    if (i == 0) v = fe_to_mont(ld_vec(p + (n - 1 - off)));
    else if (i >= k + 2) v = fe_to_mont(ld_vec(p + (2 * l - off - 1 + (i - k - 2) * l)));
    st_vec(out + t, v);
}
void launch_toeplitz_coeffs_strided(const uint64_t* polys, Fr* out, size_t n, size_t chunk_len, size_t batch,
                                    cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    size_t total = 2 * (n / chunk_len) * chunk_len * batch;
    k_toeplitz_coeffs_strided<<<grid_for(total, 256), 256, 0, st>>>(polys, out, n, chunk_len, batch); g_launch_count++;
}

// ------------------------------------------------------------------------------ blob validation
__global__ void k_fr_check_canonical(const Fr* __restrict__ vals, size_t n, size_t batch, uint32_t* ok) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    Fr a = ld_vec(vals + t);
    uint32_t cf = 0, d;
    d = sub_cc(a.l[0], FrParams::mod(0), cf);
#pragma unroll
    for (int i = 1; i < 8; i++) d = subc_cc(a.l[i], FrParams::mod(i), cf);
    (void)d;
    if (subc(0u, 0u, cf) == 0) atomicAnd(ok + t / n, 0u);   // no borrow: a >= r
}
void launch_fr_check_canonical(const uint64_t* vals, size_t n, size_t batch, uint32_t* ok, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    size_t total = n * batch;
    if (!total) return;
    k_fr_check_canonical<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const Fr*>(vals), n, batch, ok);
    g_launch_count++;
}

// ------------------------------------------------------------------------------ evaluation-form proofs
// eth.ComputeKZGProof (eth/helpers.go:179-203) with bls.EvaluatePolyInEvaluationForm (bls/globals.go:106-153)
// for a batch of polynomials given by their evaluations on the bit-reversed domain D[i] = w^brp(i)
// (eth/globals.go:60-67), one challenge z per polynomial:
//     y    = (z^n - 1) / n * sum_i f_i D_i / (z - D_i)
//     q_i  = (f_i - y) / (D_i - z)
// Pass 1 inverts the n denominators (Montgomery's trick, EVAL_CHUNK per lane) and leaves a partial sum
// per lane; pass 2 folds the partial sums into y; pass 3 forms the quotient (canonical, ready for the MSM).
#define EVAL_CHUNK 16
// domain element of position i: reverse bit order (eth DomainFr, eth/globals.go:60-67) or natural order
__device__ __forceinline__ size_t eval_domain_index(size_t i, unsigned logn, size_t stride, int bitrev) {
    return (bitrev ? (size_t)(logn ? __brev((uint32_t)i) >> (32 - logn) : 0u) : i) * stride;
}
// ch = min(n, EVAL_CHUNK) elements per lane
__global__ void __launch_bounds__(128) k_eval_form_pass1(const Fr* __restrict__ f_canon, const Fr* __restrict__ z_canon,
                                                         const Fr* __restrict__ expanded, size_t stride, size_t n, unsigned logn,
                                                         size_t batch, int bitrev, int ch, Fr* inv_den, Fr* partial, uint32_t* ok) {
    const size_t chunks = n / ch;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= chunks * batch) return;
    size_t b = t / chunks, lo = (t % chunks) * ch;
    Fr z = fe_to_mont(ld_vec(z_canon + b));
    Fr pre[EVAL_CHUNK];
    Fr acc = Fr::one();
    bool hit = false;
    for (int k = 0; k < ch; k++) {
        Fr d = fe_sub(ld_vec(expanded + eval_domain_index(lo + k, logn, stride, bitrev)), z);    // D_i - z
        pre[k] = acc;
        if (d.is_zero()) hit = true; else acc = fe_mul(acc, d);
    }
    if (hit && ok) atomicAnd(ok + b, 0u);                                                 // "invalid z challenge"
    acc = fe_inv(acc);
    Fr sum = Fr::zero();
    for (int k = ch - 1; k >= 0; k--) {
        Fr dom = ld_vec(expanded + eval_domain_index(lo + k, logn, stride, bitrev));
        Fr d = fe_sub(dom, z);
        Fr inv = Fr::zero();
        if (!d.is_zero()) { inv = fe_mul(acc, pre[k]); acc = fe_mul(acc, d); }
        st_vec(inv_den + b * n + lo + k, inv);                                            // 1 / (D_i - z), Montgomery
        Fr fi = fe_to_mont(ld_vec(f_canon + b * n + lo + k));
        sum = fe_sub(sum, fe_mul(fe_mul(fi, dom), inv));                                  // + f_i D_i / (z - D_i)
    }
    st_vec(partial + t, sum);
}
// y[b] = (z^n - 1) / n * sum of the partial sums; one block per polynomial
__global__ void __launch_bounds__(256) k_eval_form_pass2(const Fr* __restrict__ partial, size_t chunks, const Fr* __restrict__ z_canon,
                                                         unsigned logn, Fr inv_n, Fr* y_mont, Fr* y_canon) {
    __shared__ Fr sh[256];
    size_t b = blockIdx.x;
    Fr s = Fr::zero();
    for (size_t i = threadIdx.x; i < chunks; i += blockDim.x) s = fe_add(s, ld_vec(partial + b * chunks + i));
    sh[threadIdx.x] = s;
    __syncthreads();
    for (unsigned w = blockDim.x / 2; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] = fe_add(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Fr zn = fe_to_mont(ld_vec(z_canon + b));
        for (unsigned k = 0; k < logn; k++) zn = fe_mul(zn, zn);                          // z^n, n = 2^logn
        Fr y = fe_mul(fe_mul(fe_sub(zn, Fr::one()), inv_n), sh[0]);
        st_vec(y_mont + b, y);
        if (y_canon) st_vec(y_canon + b, fe_from_mont(y));
    }
}
__global__ void k_eval_form_pass3(const Fr* __restrict__ f_canon, const Fr* __restrict__ inv_den, const Fr* __restrict__ y_mont,
                                  size_t n, size_t batch, Fr* q_canon) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    Fr fi = fe_to_mont(ld_vec(f_canon + t));
    Fr q = fe_mul(fe_sub(fi, ld_vec(y_mont + t / n)), ld_vec(inv_den + t));
    st_vec(q_canon + t, fe_from_mont(q));
}
void launch_eval_form_quotient(const FrDomain& dom, const uint64_t* f_canon, const uint64_t* z_canon, unsigned logn, size_t batch, int bitrev,
                               const Fr& inv_n, Fr* inv_den /* batch n */, Fr* partial /* batch n / 16 */, Fr* y_mont /* batch */,
                               uint64_t* y_canon_or_null, uint64_t* q_canon /* batch n, or null: evaluation only */, uint32_t* ok, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    const size_t n = (size_t)1 << logn;
    const int ch = n < EVAL_CHUNK ? (int)n : EVAL_CHUNK;
    const size_t chunks = n / ch;
    if (!batch) return;
    const Fr* f = reinterpret_cast<const Fr*>(f_canon);
    const Fr* z = reinterpret_cast<const Fr*>(z_canon);
    k_eval_form_pass1<<<grid_for(chunks * batch, 128), 128, 0, st>>>(f, z, dom.expanded, dom.max_width >> logn, n, logn, batch, bitrev, ch, inv_den, partial, ok);
    k_eval_form_pass2<<<(unsigned)batch, 256, 0, st>>>(partial, chunks, z, logn, inv_n, y_mont, reinterpret_cast<Fr*>(y_canon_or_null));
    g_launch_count += 2;
    if (q_canon) { k_eval_form_pass3<<<grid_for(n * batch, 256), 256, 0, st>>>(f, inv_den, y_mont, n, batch, reinterpret_cast<Fr*>(q_canon)); g_launch_count++; }
}

// kzg_multi_proofs.go:57-71 for a batch of samples: coefficient i of sample b is divided by x_b^i (InvModFr(0) = 0, so x = 0
// clears every coefficient but the first, as in the reference), and x_b^n is returned for the caller's [x^n]_2.
// coeffs: [batch][n] Montgomery in, canonical out; x canonical; n = 2^logn.
__global__ void k_unscale_coset(Fr* coeffs, const Fr* __restrict__ x_canon, size_t n, unsigned logn, size_t batch, Fr* x_pow_n_canon) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    const size_t b = t / n, i = t % n;
    const Fr x = fe_to_mont(ld_vec(x_canon + b));
    const Fr xi = fe_inv(x);                                   // 0 -> 0
    Fr p = Fr::one(), sq = xi;                                 // xi^i by square and multiply
    for (size_t e = i; e; e >>= 1) { if (e & 1) p = fe_mul(p, sq); sq = fe_mul(sq, sq); }
    st_vec(coeffs + t, fe_from_mont(fe_mul(ld_vec(coeffs + t), p)));
    if (i == 0) {
        Fr xn = x;
        for (unsigned k = 0; k < logn; k++) xn = fe_mul(xn, xn);
        st_vec(x_pow_n_canon + b, fe_from_mont(xn));
    }
}
void launch_unscale_coset(Fr* coeffs, const uint64_t* x_canon, unsigned logn, size_t batch, uint64_t* x_pow_n_canon, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    const size_t n = (size_t)1 << logn;
    if (!batch) return;
    k_unscale_coset<<<grid_for(n * batch, 128), 128, 0, st>>>(coeffs, reinterpret_cast<const Fr*>(x_canon), n, logn, batch, reinterpret_cast<Fr*>(x_pow_n_canon));
    g_launch_count++;
}

}  // namespace b200
