// kernels_g1.cu -- G1 kernels: ABI conversion, radix-2 butterfly stages of the G1 FFT
// (fft_g1.go:33-56 _fftG1, one MulG1 + AddG1 + SubG1 per butterfly), batched scalar
// multiplication (fk20_single.go:72-74), tree folding for LinCombG1 (bls/bls_kilic.go:132-150).
//
// All of these are bound by the integer multiply-add pipe, not by HBM: one butterfly moves
// 4 x 144 B and executes ~1700 Fp multiplications (~5 x 10^5 IMADs).  Memory layout is therefore
// plain AoS Jacobian (144 B, 16 B vector accesses); the design effort goes into keeping warps
// convergent (lanes = blobs share one twiddle program) and the register footprint bounded
// (out-of-line point operations, tables in local memory).
#include "g1_dev.cuh"
#include "kernels.h"
#include <atomic>
#include "quad.cuh"

namespace b200 {

thread_local unsigned long long g_launch_count = 0;

// Launch shape of the integer-pipe bound G1 kernels.  128 registers per thread hold the working
// set of the out-of-line point operations without spills (ptxas -v), which lets 4 CTAs of 128
// threads share an SM: the IMAD.WIDE carry chains of one warp are latency bound, so issue slots
// are filled by warp count, not by ILP.
#ifndef B200_G1_BLOCK
#define B200_G1_BLOCK 128
#endif
#ifndef B200_G1_MINB
#define B200_G1_MINB 4
#endif
// Batched transforms map (butterfly, blob) to threads with the blob index fastest.  From 16 blobs on the
// lane count per butterfly is rounded up to whole warps (the padding lanes idle), so that every warp works on
// ONE butterfly and the sparse (width-5 NAF) twiddle programs stay convergent for any batch size.
__host__ __device__ inline size_t g1_lanes_for_batch(size_t batch) { return batch >= 16 ? (batch + 31) / 32 * 32 : batch; }
constexpr unsigned G1_BLOCK = B200_G1_BLOCK;
constexpr unsigned G1_MINB = B200_G1_MINB;

static inline unsigned grid_for(size_t total, unsigned block) { return (unsigned)((total + block - 1) / block); }

__device__ __forceinline__ uint32_t bitrev_u32(uint32_t v, unsigned logn) { return logn ? (__brev(v) >> (32 - logn)) : 0u; }

// ------------------------------------------------------------------------------ conversions
__global__ void k_g1_from_abi(const uint64_t* __restrict__ in, G1J* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1J p = ld_vec(reinterpret_cast<const G1J*>(in) + i);
    if (p.is_inf()) { st_vec(out + i, G1J::infinity()); return; }
    p.x = fe_to_mont(p.x); p.y = fe_to_mont(p.y); p.z = fe_to_mont(p.z);
    st_vec(out + i, p);
}
__global__ void k_g1_to_abi(const G1J* __restrict__ in, uint64_t* __restrict__ out, size_t n, size_t batch,
                            size_t estride, size_t bstride, int bitrev, unsigned logn) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    size_t b = t / n, i = t % n;
    size_t src = bitrev ? bitrev_u32((uint32_t)i, logn) : i;
    G1J p = ld_vec(in + b * bstride + src * estride);
    if (p.is_inf()) p = G1J::infinity();
    else { p.x = fe_from_mont(p.x); p.y = fe_from_mont(p.y); p.z = fe_from_mont(p.z); }
    st_vec(reinterpret_cast<G1J*>(out) + t, p);
}
__global__ void k_g1_fill_infinity(G1J* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_vec(p + i, G1J::infinity());
}

void launch_g1_from_abi(const uint64_t* in, G1J* out, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_g1_from_abi<<<grid_for(n, 128), 128, 0, st>>>(in, out, n); g_launch_count++;
}
void launch_g1_to_abi(const G1J* in, uint64_t* out, size_t n, size_t batch, size_t estride, size_t bstride, int bitrev,
                      unsigned logn, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_g1_to_abi<<<grid_for(n * batch, 128), 128, 0, st>>>(in, out, n, batch, estride, bstride, bitrev, logn); g_launch_count++;
}
void launch_g1_fill_infinity(G1J* p, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_g1_fill_infinity<<<grid_for(n, 256), 256, 0, st>>>(p, n); g_launch_count++;
}

// ------------------------------------------------------------------------------ compression
// ToCompressedG1 (bls/bls_kilic.go:114-116; ZCash 48-byte form: big-endian affine x, bit 7
// "compressed", bit 6 infinity, bit 5 y > (p-1)/2) for whole arrays on the device: the step right
// after the hot path (proofs and commitments leave the system compressed) and a third of the
// device->host bytes.  One thread normalises COMP_GROUP points with a single Fermat inversion
// (Montgomery's trick on the Z coordinates): ~45 Fp products per point instead of ~580.
#define COMP_GROUP 16
struct alignas(16) Comp48 { uint32_t w[12]; };
__device__ __forceinline__ bool fp_canon_gt_half_dev(const Fp& y) {   // y > (p-1)/2, canonical limbs
    constexpr uint32_t half[12] = B200_FP_HALF;
    uint32_t cf = 0, t;
    t = sub_cc(half[0], y.l[0], cf);
#pragma unroll
    for (int i = 1; i < 12; i++) t = subc_cc(half[i], y.l[i], cf);
    (void)t;
    return subc(0u, 0u, cf) != 0;   // borrow <=> half < y
}
static __device__ __noinline__ Fp fp_inv_fermat(const Fp* a) {         // a^(p-2); a != 0
    constexpr uint32_t e[12] = B200_FP_MODM2;
    Fp acc = *a;
    for (int i = 379; i >= 0; i--) {                                  // bit 380 is the top bit of p - 2
        acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = fp_mul(acc, *a);
    }
    return acc;
}
__device__ __forceinline__ size_t comp_src_index(size_t t, size_t n, size_t estride, size_t bstride, int bitrev, unsigned logn) {
    size_t b = t / n, i = t % n;
    size_t src = bitrev ? bitrev_u32((uint32_t)i, logn) : i;
    return b * bstride + src * estride;
}
__global__ void __launch_bounds__(128) k_g1_compress(const G1J* __restrict__ in, Comp48* __restrict__ out, size_t n, size_t batch,
                                                     size_t estride, size_t bstride, int bitrev, unsigned logn) {
    const size_t total = n * batch;
    const size_t t0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * COMP_GROUP;
    if (t0 >= total) return;
    const int cnt = (int)(total - t0 < COMP_GROUP ? total - t0 : COMP_GROUP);
    Fp pre[COMP_GROUP];
    Fp acc = Fp::one();
    for (int k = 0; k < cnt; k++) {
        Fp z = ld_vec(&in[comp_src_index(t0 + k, n, estride, bstride, bitrev, logn)].z);
        pre[k] = acc;
        if (!z.is_zero()) acc = fp_mul(acc, z);
    }
    Fp inv = fp_inv_fermat(&acc);
    Fp raw_one = Fp::zero(); raw_one.l[0] = 1;                        // x * 1 / R: leaves Montgomery form
    for (int k = cnt - 1; k >= 0; k--) {
        G1J p = ld_vec(&in[comp_src_index(t0 + k, n, estride, bstride, bitrev, logn)]);
        Comp48 c;
        if (p.is_inf()) {
#pragma unroll
            for (int j = 0; j < 12; j++) c.w[j] = 0;
            c.w[0] = 0xC0u;
        } else {
            Fp zi = fp_mul(inv, pre[k]);
            inv = fp_mul(inv, p.z);
            Fp zi2 = fp_sqr(zi);
            Fp x = fp_mul(fp_mul(p.x, zi2), raw_one);
            Fp y = fp_mul(fp_mul(p.y, fp_mul(zi2, zi)), raw_one);
#pragma unroll
            for (int j = 0; j < 12; j++) c.w[j] = __byte_perm(x.l[11 - j], 0u, 0x0123);   // big-endian bytes
            c.w[0] |= 0x80u | (fp_canon_gt_half_dev(y) ? 0x20u : 0u);
        }
        out[t0 + k] = c;
    }
}
void launch_g1_compress(const G1J* in, uint8_t* out48, size_t n, size_t batch, size_t estride, size_t bstride, int bitrev,
                        unsigned logn, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    size_t groups = (n * batch + COMP_GROUP - 1) / COMP_GROUP;
    k_g1_compress<<<grid_for(groups, 128), 128, 0, st>>>(in, reinterpret_cast<Comp48*>(out48), n, batch, estride, bstride, bitrev, logn);
    g_launch_count++;
}

// ------------------------------------------------------------------------------ decompression
// FromCompressedG1 (bls/bls_kilic.go:118-121 -> kilic FromCompressed) for whole arrays: setup files hold thousands to
// millions of 48-byte points and every one costs a square root (a 381-bit exponentiation) and a subgroup check.
// One point per thread: flags, x < p, y = (x^3 + 4)^((p + 1) / 4) with y^2 verified, sign from bit 5, then the
// subgroup test [z^2] P == (beta x, -y) (g1.cuh: g1_in_subgroup; a 128-bit multiplication).
__global__ void __launch_bounds__(128) k_g1_decompress(const Comp48* __restrict__ in, G1J* __restrict__ out_abi, uint32_t* __restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Comp48 c = in[i];
    const uint32_t b0 = c.w[0] & 0xffu;
    G1J res = G1J::infinity();                                        // all-zero limbs: infinity in the ABI encoding too
    uint32_t rc = 0;
    if (!(b0 & 0x80u)) rc = 1;
    else if (b0 & 0x40u) {
        uint32_t rest = c.w[0] & 0xffffff3fu;                         // every other bit of an infinity encoding must be clear
#pragma unroll
        for (int j = 1; j < 12; j++) rest |= c.w[j];
        if (rest) rc = 1;
    } else {
        Fp x;
#pragma unroll
        for (int j = 0; j < 12; j++) x.l[11 - j] = __byte_perm(c.w[j], 0u, 0x0123);
        x.l[11] &= 0x1fffffffu;
        uint32_t cf = 0, t;
        t = sub_cc(x.l[0], FpParams::mod(0), cf);
#pragma unroll
        for (int j = 1; j < 12; j++) t = subc_cc(x.l[j], FpParams::mod(j), cf);
        (void)t;
        if (subc(0u, 0u, cf) == 0) rc = 1;                            // no borrow: x >= p
        else {
            const Fp xm = fp_mul(x, Fp::r2());
            const Fp rhs = fe_add(fp_mul(fp_sqr(xm), xm), fp_const_four());
            constexpr uint32_t e[12] = B200_FP_SQRT_EXP;               // (p + 1) / 4, 379 bits
            Fp y = rhs;
            int top = 383;
            while (top > 0 && !((e[top >> 5] >> (top & 31)) & 1u)) top--;
            for (int bit = top - 1; bit >= 0; bit--) {
                y = fp_sqr(y);
                if ((e[bit >> 5] >> (bit & 31)) & 1u) y = fp_mul(y, rhs);
            }
            if (fp_sqr(y) != rhs) rc = 2;
            else {
                Fp raw_one = Fp::zero(); raw_one.l[0] = 1;
                Fp yc = fp_mul(y, raw_one);
                if (fp_canon_gt_half_dev(yc) != ((b0 & 0x20u) != 0)) { y = fe_neg(y); yc = fp_mul(y, raw_one); }
                G1A pa; pa.x = xm; pa.y = y;
                G1J acc; acc.x = xm; acc.y = y; acc.z = Fp::one();
                constexpr uint32_t z2[4] = B200_GLV_Z2;               // top bit is bit 127
                for (int bit = 126; bit >= 0; bit--) {
                    g1_dbl_ni(&acc, &acc);
                    if ((z2[bit >> 5] >> (bit & 31)) & 1u) g1_add_mixed_ni(&acc, &acc, &pa);
                }
                G1J pj; pj.x = xm; pj.y = y; pj.z = Fp::one();
                if (!g1_equal(acc, g1_endo(pj))) rc = 3;
                else { res.x = x; res.y = yc; res.z = Fp::zero(); res.z.l[0] = 1; }
            }
        }
    }
    st_vec(out_abi + i, res);
    status[i] = rc;
}
void launch_g1_decompress(const uint8_t* in48, uint64_t* out_abi, uint32_t* status, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_g1_decompress<<<grid_for(n, 128), 128, 0, st>>>(reinterpret_cast<const Comp48*>(in48), reinterpret_cast<G1J*>(out_abi), status, n);
    g_launch_count++;
}

// ------------------------------------------------------------------------------ FFT stage
// thread <-> (butterfly q, blob b) with b fastest: when batch is a multiple of 32 every lane of a
// warp runs the same twiddle program on a different blob, so the digit branches are uniform.
// across_blocks (few transforms, at least 32 blocks in this stage): lanes run over the BLOCKS of the stage instead, so a
// warp again holds 32 butterflies with one and the same twiddle w^j and can use the sparse (width-5 NAF) programs.
// B200_STAGE_SMEM_TABLE (experiment, off by default): the 8-entry table of the twiddle product lives in shared memory
// (16 coordinates x 48 B per thread = 96 KiB per CTA, hence 2 CTAs per SM) instead of the thread's stack.
#ifdef B200_STAGE_SMEM_TABLE
#define STAGE_MINB 2
#define STAGE_SMEM_BYTES (16u * 48u * G1_BLOCK)
#define STAGE_EXT reinterpret_cast<Fp*>(stage_smem) + threadIdx.x, G1_BLOCK
#else
#define STAGE_MINB G1_MINB
#define STAGE_SMEM_BYTES 0u
#define STAGE_EXT nullptr, 1u
#endif
template <bool DIF>
__global__ void __launch_bounds__(G1_BLOCK, STAGE_MINB) k_g1_fft_stage(G1J* data, size_t n_half, size_t batch, size_t m, size_t estride,
                                                      size_t bstride, const ScalarProgram* __restrict__ progs,
                                                      size_t prog_stride, int across_blocks, size_t jmul, size_t bmul, size_t joff) {
#ifdef B200_STAGE_SMEM_TABLE
    extern __shared__ uint4 stage_smem[];
#endif
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t b, q, j;
    if (across_blocks) {
        const size_t nblocks = n_half / m;                     // multiple of 32
        if (t >= n_half * batch) return;
        const size_t blk = t % nblocks, r = t / nblocks;
        b = r % batch; j = r / batch;
        q = blk * m + j;
    } else {
        // lanes of a butterfly: the batch, padded to whole warps (idle lanes) so that a warp never mixes twiddles
        const size_t lanes = g1_lanes_for_batch(batch);
        if (t >= n_half * lanes) return;
        b = t % lanes; q = t / lanes;
        if (b >= batch) return;
        j = q & (m - 1);
    }
    size_t i0 = 2 * q - j, i1 = i0 + m;
    G1J* p0 = data + b * bstride + i0 * estride;
    G1J* p1 = data + b * bstride + i1 * estride;
    // twiddle of the butterfly: position j inside the block; the strided sub-transforms of a sharded merge (api.cu) add the
    // element's position in its sub-range (b) and the sub-range's offset: index (j jmul + b bmul + joff)
    const ScalarProgram* prog = progs + (j * jmul + b * bmul + joff) * prog_stride;
    G1J x0 = ld_vec(p0), x1 = ld_vec(p1), s, d;
    if (DIF) {
        g1_add_sub_ni(&s, &d, &x0, &x1);
        g1_mul_program(&x1, &d, prog, STAGE_EXT);
        st_vec(p0, s);
        st_vec(p1, x1);
    } else {
        g1_mul_program(&d, &x1, prog, STAGE_EXT);
        g1_add_sub_ni(&s, &x1, &x0, &d);
        st_vec(p0, s);
        st_vec(p1, x1);
    }
}
// The same stage with one butterfly per QUAD of lanes (quad.cuh), for launches too small to occupy the chip (one or a few
// transforms): the butterfly's chain of ~1450 dependent products becomes ~650-750 dependent levels.  Lane mapping as above on
// the quad index; with across_blocks the 8 quads of a warp share the twiddle (sparse program), otherwise fixed windows.
template <bool DIF>
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_fft_stage_quad(G1J* data, size_t n_half, size_t batch, size_t m, size_t estride,
                                                                         size_t bstride, const ScalarProgram* __restrict__ progs,
                                                                         size_t prog_stride, int across_blocks, size_t jmul, size_t bmul, size_t joff) {
    const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool valid = t < n_half * batch;             // quads past the end idle through the collective operations
    const size_t tc = valid ? t : 0;
    size_t b, q, j;
    if (across_blocks) {
        const size_t nblocks = n_half / m;             // multiple of 8: the quads of a warp share (b, j)
        const size_t blk = tc % nblocks, r = tc / nblocks;
        b = r % batch; j = r / batch;
        q = blk * m + j;
    } else {
        b = tc % batch; q = tc / batch;
        j = q & (m - 1);
    }
    const size_t i0 = 2 * q - j, i1 = i0 + m;
    G1J* p0 = data + b * bstride + i0 * estride;
    G1J* p1 = data + b * bstride + i1 * estride;
    const ScalarProgram* prog = progs + (j * jmul + b * bmul + joff) * prog_stride;
    G1J x0 = G1J::infinity(), x1 = G1J::infinity(), s, d;
    if (valid) { x0 = ld_vec(p0); x1 = ld_vec(p1); }
    if (DIF) {
        quad_add_sub(&s, &d, &x0, &x1);
        quad_mul_program(&x1, &d, prog);
    } else {
        quad_mul_program(&d, &x1, prog);
        quad_add_sub(&s, &x1, &x0, &d);
    }
    if (valid && (threadIdx.x & 3u) == 0) { st_vec(p0, s); st_vec(p1, x1); }
}
// Butterflies per launch up to which a quad of lanes per butterfly is used (quad.cuh): a point doubling in 3 dependent
// products instead of 7, an addition in 5 instead of 11-16.  One polynomial at n = 4096 (2048 / 4096 butterflies per stage):
// FFTG1 18.8 -> 11.6 ms, FK20Single 37.4 -> 24.5 ms, DAUsingFK20 41.9 -> 26.3 ms (profiles/r02_quad_stage.txt).  It spends
// 4 x the warps and ~1.7 x the multiplier cycles per butterfly, so it only pays while the launch leaves the machine mostly
// idle: above ~2 quad warps per SM sub-partition the multiplier pipe is the limit again (hence 8192), and with several
// callers in flight the one-lane kernels give more aggregate throughput (8 concurrent callers: 103 against 80
// polynomials/s), which is what g1_set_quad_allowed carries in from the API layer, once per call.
#ifndef B200_QUAD_STAGE_MAX
#define B200_QUAD_STAGE_MAX 8192
#endif
static thread_local bool t_quad_allowed = true;     // per call (API layer: concurrency at the start of the call)
static std::atomic<bool> g_quad_enabled{true};      // process wide (b200_set_latency_mode), also seen by the device-pointer entry points
void g1_set_quad_allowed(bool allowed) { t_quad_allowed = allowed; }
void g1_set_quad_enabled(bool enabled) { g_quad_enabled = enabled; }
bool g1_stage_uses_quads(size_t n_half, size_t batch) {
    return t_quad_allowed && g_quad_enabled.load() && batch < 16 && n_half * batch <= B200_QUAD_STAGE_MAX;
}

static void stage_smem_opt_in() {
#ifdef B200_STAGE_SMEM_TABLE
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaFuncSetAttribute(k_g1_fft_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGE_SMEM_BYTES);
    cudaFuncSetAttribute(k_g1_fft_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STAGE_SMEM_BYTES);
    done[dev] = true;
#endif
}
void launch_g1_fft_stage(G1J* data, size_t n_half, size_t batch, size_t m, size_t estride, size_t bstride, bool dif,
                         const ScalarProgram* progs, size_t prog_stride, cudaStream_t st, int across_blocks, size_t jmul, size_t bmul, size_t joff) {
    ProfScope prof_scope(PROF_G1_FFT_STAGE, st);
    size_t total = across_blocks ? n_half * batch : n_half * g1_lanes_for_batch(batch);
    if (!total || !batch) return;
    if (g1_stage_uses_quads(n_half, batch)) {
        const size_t lanes = (n_half * batch * 4 + 31) / 32 * 32;       // whole warps: the quad operations are warp-collective
        if (dif) k_g1_fft_stage_quad<true><<<grid_for(lanes, G1_BLOCK), G1_BLOCK, 0, st>>>(data, n_half, batch, m, estride, bstride, progs, prog_stride, across_blocks, jmul, bmul, joff);
        else k_g1_fft_stage_quad<false><<<grid_for(lanes, G1_BLOCK), G1_BLOCK, 0, st>>>(data, n_half, batch, m, estride, bstride, progs, prog_stride, across_blocks, jmul, bmul, joff);
        g_launch_count++;
        return;
    }
    stage_smem_opt_in();
    if (dif) k_g1_fft_stage<true><<<grid_for(total, G1_BLOCK), G1_BLOCK, STAGE_SMEM_BYTES, st>>>(data, n_half, batch, m, estride, bstride, progs, prog_stride, across_blocks, jmul, bmul, joff);
    else k_g1_fft_stage<false><<<grid_for(total, G1_BLOCK), G1_BLOCK, STAGE_SMEM_BYTES, st>>>(data, n_half, batch, m, estride, bstride, progs, prog_stride, across_blocks, jmul, bmul, joff);
    g_launch_count++;
}

// One decimation-in-frequency stage of a single transform, keeping only the half of the outputs that a
// rank owning one of the 2^s output blocks needs: out[i] = in[i] + in[i + m] (upper block) or
// w^i (in[i] - in[i + m]) (lower block), i < m.  s such stages take the all-gathered input of size n down
// to the rank's block of size n / 2^s with 1/2 + 1/4 + .. of the butterflies instead of s n / 2.
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_dif_half_stage(const G1J* __restrict__ in, G1J* __restrict__ out, size_t m, int lower,
                                                                         const ScalarProgram* __restrict__ progs, size_t prog_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    G1J x0 = ld_vec(in + i), x1 = ld_vec(in + i + m), s, d;
    g1_add_sub_ni(&s, &d, &x0, &x1);
    if (lower) { g1_mul_program(&x1, &d, progs + i * prog_stride); st_vec(out + i, x1); }
    else st_vec(out + i, s);
}
void launch_g1_dif_half_stage(const G1J* in, G1J* out, size_t m, int lower, const ScalarProgram* progs, size_t prog_stride, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_FFT_STAGE, st);
    if (!m) return;
    k_g1_dif_half_stage<<<grid_for(m, G1_BLOCK), G1_BLOCK, 0, st>>>(in, out, m, lower, progs, prog_stride);
    g_launch_count++;
}

// ------------------------------------------------------------------------------ scalar muls
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_mul_var(const G1J* pts, size_t pts_bstride,
                                                    const Fr* __restrict__ k, int k_is_mont, G1J* out,
                                                    size_t out_bstride, size_t n, size_t batch) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    // blob fastest, so that (for shared bases) a warp reads one base point
    size_t b = t % batch, i = t / batch;
    G1J p = ld_vec(pts + b * pts_bstride + i);
    Fr s = ld_vec(k + b * n + i);
    if (k_is_mont) s = fe_from_mont(s);
    G1J r;
    g1_mul_var(&r, &p, s.l);
    st_vec(out + b * out_bstride + i, r);
}
void launch_g1_mul_var(const G1J* pts, size_t pts_bstride, const Fr* k, int k_is_mont, G1J* out, size_t out_bstride,
                       size_t n, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_MUL, st);
    if (!n || !batch) return;
    k_g1_mul_var<<<grid_for(n * batch, G1_BLOCK), G1_BLOCK, 0, st>>>(pts, pts_bstride, k, k_is_mont, out, out_bstride, n, batch);
    g_launch_count++;
}

__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_mul_programs(G1J* data, size_t n, size_t batch, size_t estride, size_t bstride,
                                                         const ScalarProgram* __restrict__ progs, size_t prog_stride,
                                                         int bitrev, unsigned logn) {
    const size_t lanes = g1_lanes_for_batch(batch);
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * lanes) return;
    size_t b = t % lanes, i = t / lanes;
    if (b >= batch) return;
    size_t pi = bitrev ? bitrev_u32((uint32_t)i, logn) : i;
    G1J* p = data + b * bstride + i * estride;
    G1J x = ld_vec(p), r;
    g1_mul_program(&r, &x, progs + pi * prog_stride);
    st_vec(p, r);
}
// the same with a quad of lanes per product (latency mode, small launches: see g1_stage_uses_quads)
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_mul_programs_quad(G1J* data, size_t n, size_t batch, size_t estride, size_t bstride,
                                                                            const ScalarProgram* __restrict__ progs, size_t prog_stride,
                                                                            int bitrev, unsigned logn) {
    const size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool valid = t < n * batch;                  // quads past the end idle through the collective operations
    const size_t tc = valid ? t : 0;
    const size_t b = tc % batch, i = tc / batch;
    const size_t pi = bitrev ? bitrev_u32((uint32_t)i, logn) : i;
    G1J* p = data + b * bstride + i * estride;
    G1J x = G1J::infinity(), r;
    if (valid) x = ld_vec(p);
    quad_mul_program(&r, &x, progs + pi * prog_stride);
    if (valid && (threadIdx.x & 3u) == 0) st_vec(p, r);
}
void launch_g1_mul_programs(G1J* data, size_t n, size_t batch, size_t estride, size_t bstride, const ScalarProgram* progs,
                            size_t prog_stride, int bitrev, unsigned logn, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_MUL, st);
    if (!n || !batch) return;
    if (g1_stage_uses_quads(n, batch)) {
        const size_t lanes = (n * batch * 4 + 31) / 32 * 32;
        k_g1_mul_programs_quad<<<grid_for(lanes, G1_BLOCK), G1_BLOCK, 0, st>>>(data, n, batch, estride, bstride, progs, prog_stride, bitrev, logn);
    } else {
        k_g1_mul_programs<<<grid_for(n * g1_lanes_for_batch(batch), G1_BLOCK), G1_BLOCK, 0, st>>>(data, n, batch, estride, bstride, progs, prog_stride, bitrev, logn);
    }
    g_launch_count++;
}

// ------------------------------------------------------------------------------ fixed-base tables
// The bases of ToeplitzPart2 (xExtFFT, fk20_single.go:72-74) and of CommitToPoly (SecretG1,
// kzg_single_proofs.go:17-19) are fixed per settings object, so their scalar multiplications
// become table look-ups: signed W-bit windows, table[i][w][d-1] = d * 2^(W w) * P_i in affine
// form.  W = 8: 32 windows x 128 entries x 96 B = 384 KiB per base (3 GiB for the 8192 bases of
// the n = 4096 FK20 settings -- HBM is what a B200 has plenty of), one product = 32 mixed
// additions (~350 Fp multiplications) instead of ~1900.  W = 12: 22 windows x 2048 entries = 4.1 MiB per
// base and 22 mixed additions, used while the table stays small against HBM (the n = 4096 settings:
// 17 + 35 GB).  W = 4: 64 windows x 8 entries = 48 KiB
// per base and 64 mixed additions, for settings whose W = 8 table would not fit (config 5: 2.1 M
// bases).
static inline unsigned fb_windows(int W) { return (unsigned)((256 + W - 1) / W); }
static inline unsigned fb_entries(int W) { return 1u << (W - 1); }
#define FB_CHUNK 16
// bases[i][w] = 2^(W w) P_i (Jacobian)
__global__ void __launch_bounds__(128) k_fb_bases(const G1J* __restrict__ pts, G1J* __restrict__ bases, size_t n, int W, unsigned nw) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1J p = ld_vec(pts + i);
    for (unsigned w = 0; w < nw; w++) {
        st_vec(bases + i * nw + w, p);
        if (w + 1 < nw) for (int k = 0; k < W; k++) g1_dbl_ni(&p, &p);
    }
}
// One thread fills FB_CHUNK consecutive entries of one (base, window) row: start = (c CH + 1) B by double-and-add,
// then a chain of additions of B, then ONE inversion for the whole chunk (Montgomery's trick on the Z coordinates).
// ~70 Fp products per entry instead of ~700 when every entry is built and inverted on its own.
__global__ void __launch_bounds__(128) k_fb_entries(const G1J* __restrict__ bases, G1A* __restrict__ table, size_t n_rows, unsigned D) {
    const unsigned CH = D < FB_CHUNK ? D : FB_CHUNK, chunks = D / CH;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * chunks) return;
    const size_t row = t / chunks;
    const unsigned first = (unsigned)(t % chunks) * CH + 1;          // multiple of the first entry of this chunk
    G1J b = ld_vec(bases + row), acc = G1J::infinity();
    for (int bit = 11; bit >= 0; bit--) {                             // first <= 2^11
        if (!acc.is_inf()) g1_dbl_ni(&acc, &acc);
        if ((first >> bit) & 1u) g1_add_ni(&acc, &acc, &b);
    }
    G1J pts[FB_CHUNK];
    Fp pre[FB_CHUNK];
    Fp prod = Fp::one();
    for (unsigned e = 0; e < CH; e++) {
        pts[e] = acc;
        pre[e] = prod;
        if (!acc.is_inf()) prod = fp_mul(prod, acc.z);
        if (e + 1 < CH) g1_add_ni(&acc, &acc, &b);
    }
    Fp inv = fp_inv_fermat(&prod);
    G1A* out = table + row * D + (first - 1);
    for (int e = (int)CH - 1; e >= 0; e--) {
        G1A a;
        if (pts[e].is_inf()) { a.x = Fp::zero(); a.y = Fp::zero(); }
        else {
            Fp zi = fp_mul(inv, pre[e]);
            inv = fp_mul(inv, pts[e].z);
            Fp zi2 = fp_sqr(zi);
            a.x = fp_mul(pts[e].x, zi2);
            a.y = fp_mul(pts[e].y, fp_mul(zi2, zi));
        }
        st_vec(out + e, a);
    }
}
void launch_fixed_base_table(const G1J* pts, size_t n, G1J* bases_tmp, G1A* table, int W, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    const unsigned nw = fb_windows(W), D = fb_entries(W), CH = D < FB_CHUNK ? D : FB_CHUNK;
    k_fb_bases<<<grid_for(n, 128), 128, 0, st>>>(pts, bases_tmp, n, W, nw);
    k_fb_entries<<<grid_for(n * nw * (D / CH), 128), 128, 0, st>>>(bases_tmp, table, n * nw, D);
    g_launch_count += 2;
}
// out[b * out_bstride + i] = k[b * n + i] * P_i through the table (thread <-> (i, blob), blob fastest:
// a warp walks the same table row)
// acc += s * P through the window row of P (s canonical)
template <int W>
__device__ __forceinline__ void fb_accumulate(G1J* acc, const G1A* __restrict__ row, const Fr& s) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1), MASK = (1u << W) - 1u;
    unsigned carry = 0;
    for (unsigned w = 0; w < NW; w++) {
        const unsigned off = w * W, li = off >> 5, sh = off & 31u;      // window bits may straddle two limbs (W = 12)
        unsigned bits = s.l[li] >> sh;
        if (sh + W > 32 && li + 1 < 8) bits |= s.l[li + 1] << (32 - sh);
        unsigned d = (bits & MASK) + carry;
        bool neg = d > D;
        if (neg) { d = (1u << W) - d; carry = 1; } else carry = 0;
        if (d) {
            G1A p = ld_vec(row + w * D + (d - 1));
            if (neg) p.y = fe_neg(p.y);
            g1_add_mixed_ni(acc, acc, &p);
        }
    }
}
// Latency form of the look-ups (small launches): the windows of one product are spread over FB_SPLIT adjacent lanes, each
// lane adds its share (windows w = lane, lane + FB_SPLIT, ..), and the partial sums are combined with a shuffle tree:
// 32 dependent additions become 4 + 3 tree levels (W = 8).  The digit recoding carries from window to window, so every lane
// walks all the windows and adds only at its own.
#define FB_SPLIT 8
template <int W>
__device__ __forceinline__ void fb_accumulate_strided(G1J* acc, const G1A* __restrict__ row, const Fr& s, unsigned first) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1), MASK = (1u << W) - 1u, PER = (NW + FB_SPLIT - 1) / FB_SPLIT;
    // pass 1: the signed digits of this lane's windows (the recoding carries from window to window: all of them are walked)
    int mine[PER];
#pragma unroll
    for (unsigned j = 0; j < PER; j++) mine[j] = 0;
    unsigned carry = 0;
#pragma unroll
    for (unsigned w = 0; w < NW; w++) {
        const unsigned off = w * W, li = off >> 5, sh = off & 31u;
        unsigned bits = s.l[li] >> sh;
        if (sh + W > 32 && li + 1 < 8) bits |= s.l[li + 1] << (32 - sh);
        unsigned d = (bits & MASK) + carry;
        const bool neg = d > D;
        if (neg) { d = (1u << W) - d; carry = 1; } else carry = 0;
        if ((w % FB_SPLIT) == first) mine[w / FB_SPLIT] = neg ? -(int)d : (int)d;
    }
    // pass 2: the j-th addition of every lane in the same iteration (lanes that diverge serialise: one addition per window
    // and warp would cost what the unsplit loop costs)
#pragma unroll
    for (unsigned j = 0; j < PER; j++) {
        const int dg = mine[j];
        const unsigned w = j * FB_SPLIT + first;
        if (dg != 0 && w < NW) {
            const unsigned d = dg < 0 ? (unsigned)(-dg) : (unsigned)dg;
            G1A p = ld_vec(row + w * D + (d - 1));
            if (dg < 0) p.y = fe_neg(p.y);
            g1_add_mixed_ni(acc, acc, &p);
        }
    }
}
__device__ __forceinline__ void g1_split_sum(G1J& acc) {          // sum over FB_SPLIT adjacent lanes (whole warps only)
    for (unsigned off = FB_SPLIT >> 1; off >= 1; off >>= 1) {
        G1J other;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            other.x.l[i] = __shfl_xor_sync(0xffffffffu, acc.x.l[i], off);
            other.y.l[i] = __shfl_xor_sync(0xffffffffu, acc.y.l[i], off);
            other.z.l[i] = __shfl_xor_sync(0xffffffffu, acc.z.l[i], off);
        }
        g1_add_ni(&acc, &acc, &other);
    }
}
template <int W>
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_mul_fixed_base_split(const G1A* __restrict__ table, const Fr* __restrict__ k, int k_is_mont,
                                                                               G1J* out, size_t out_bstride, size_t n, size_t batch) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1);
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, t = tid / FB_SPLIT;
    const unsigned lane = (unsigned)(tid % FB_SPLIT);
    const bool valid = t < n * batch;                              // the grid is whole warps: idle groups walk the tree with infinity
    G1J acc = G1J::infinity();
    size_t b = 0, i = 0;
    if (valid) {
        b = t % batch; i = t / batch;
        Fr s = ld_vec(k + b * n + i);
        if (k_is_mont) s = fe_from_mont(s);
        fb_accumulate_strided<W>(&acc, table + i * (size_t)(NW * D), s, lane);
    }
    g1_split_sum(acc);
    if (valid && lane == 0) st_vec(out + b * out_bstride + i, acc);
}
template <int W>
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_g1_mul_fixed_base(const G1A* __restrict__ table, const Fr* __restrict__ k, int k_is_mont,
                                                                         G1J* out, size_t out_bstride, size_t n, size_t batch) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1);
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    size_t b = t % batch, i = t / batch;
    Fr s = ld_vec(k + b * n + i);
    if (k_is_mont) s = fe_from_mont(s);
    G1J acc = G1J::infinity();
    fb_accumulate<W>(&acc, table + i * (size_t)(NW * D), s);
    st_vec(out + b * out_bstride + i, acc);
}
// ToeplitzPart2 of FK20Single fused with the first two stages of the inverse transform over the odd slots
// (api.cu: dev_fk20).  The transform's inputs are x[m] = c[2m+1] X[2m+1] with FIXED points X, so two
// decimation-in-frequency stages only mix the scalars: for q = k / 4, p < q and x_t = x[p + t q],
//     out[p]       = x_0 + x_1 + x_2 + x_3
//     out[p + q]   = u (x_0 - x_1 + x_2 - x_3)                       u  = w_(k/2)^-p
//     out[p + 2q]  = v (x_0 - x_2) + v' (x_1 - x_3)                   v  = w_k^-p,  v' = w_k^-(p+q)
//     out[p + 3q]  = u (v (x_0 - x_2) - v' (x_1 - x_3))
// i.e. four 4-term fixed-base sums (4 x 22 mixed additions at W = 12) instead of 4 look-up products, 4 twiddle
// products (~1450 Fp products each) and 4 butterflies.  Even slots i: plain c[i] X[i].  c is Montgomery,
// rev[t rstride] = w_k^-t (Montgomery).  Threads <-> (slot, blob), blob fastest: a warp shares slot and scalars' roles.
template <int W>
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_fk20_part2_fold2(const G1A* __restrict__ table, const Fr* __restrict__ c,
                                                                        const Fr* __restrict__ rev, size_t rstride, G1J* out,
                                                                        size_t out_bstride, size_t k, size_t batch) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1);
    constexpr size_t ROW = (size_t)NW * D;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * k * batch) return;
    // the odd slots (4-term sums) come first in grid order, the cheap even slots fill the tail of the launch
    const size_t b = t % batch, ii = t / batch;
    const size_t i = ii < k ? 2 * ii + 1 : 2 * (ii - k);
    const Fr* cb = c + b * 2 * k;
    G1J acc = G1J::infinity();
    if (!(i & 1)) {
        fb_accumulate<W>(&acc, table + i * ROW, fe_from_mont(ld_vec(cb + i)));
    } else {
        const size_t q = k / 4, m = i >> 1, o = m / q, p = m % q;
        Fr u = ld_vec(rev + 2 * p * rstride), v = ld_vec(rev + p * rstride), v2 = ld_vec(rev + (p + q) * rstride);
        for (unsigned tt = 0; tt < 4; tt++) {
            const size_t src = 2 * (p + tt * q) + 1;
            Fr s = ld_vec(cb + src);
            bool neg = false;
            if (o == 1) { s = fe_mul(s, u); neg = tt & 1; }
            else if (o == 2) { s = fe_mul(s, (tt & 1) ? v2 : v); neg = tt >= 2; }
            else if (o == 3) { s = fe_mul(fe_mul(s, (tt & 1) ? v2 : v), u); neg = (tt == 1 || tt == 2); }
            if (neg) s = fe_neg(s);
            fb_accumulate<W>(&acc, table + src * ROW, fe_from_mont(s));
        }
    }
    st_vec(out + b * out_bstride + i, acc);
}
// latency form of k_fk20_part2_fold2: FB_SPLIT lanes per output slot (the 4 x 32 additions of an odd slot become 16 + 3 tree levels)
template <int W>
__global__ void __launch_bounds__(G1_BLOCK, G1_MINB) k_fk20_part2_fold2_split(const G1A* __restrict__ table, const Fr* __restrict__ c,
                                                                              const Fr* __restrict__ rev, size_t rstride, G1J* out,
                                                                              size_t out_bstride, size_t k, size_t batch) {
    constexpr unsigned NW = (256 + W - 1) / W, D = 1u << (W - 1);
    constexpr size_t ROW = (size_t)NW * D;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, t = tid / FB_SPLIT;
    const unsigned lane = (unsigned)(tid % FB_SPLIT);
    const bool valid = t < 2 * k * batch;
    G1J acc = G1J::infinity();
    size_t b = 0, i = 0;
    if (valid) {
        b = t % batch;
        const size_t ii = t / batch;
        i = ii < k ? 2 * ii + 1 : 2 * (ii - k);
        const Fr* cb = c + b * 2 * k;
        if (!(i & 1)) {
            fb_accumulate_strided<W>(&acc, table + i * ROW, fe_from_mont(ld_vec(cb + i)), lane);
        } else {
            const size_t q = k / 4, m = i >> 1, o = m / q, p = m % q;
            Fr u = ld_vec(rev + 2 * p * rstride), v = ld_vec(rev + p * rstride), v2 = ld_vec(rev + (p + q) * rstride);
            for (unsigned tt = 0; tt < 4; tt++) {
                const size_t src = 2 * (p + tt * q) + 1;
                Fr s = ld_vec(cb + src);
                bool neg = false;
                if (o == 1) { s = fe_mul(s, u); neg = tt & 1; }
                else if (o == 2) { s = fe_mul(s, (tt & 1) ? v2 : v); neg = tt >= 2; }
                else if (o == 3) { s = fe_mul(fe_mul(s, (tt & 1) ? v2 : v), u); neg = (tt == 1 || tt == 2); }
                if (neg) s = fe_neg(s);
                fb_accumulate_strided<W>(&acc, table + src * ROW, fe_from_mont(s), lane);
            }
        }
    }
    g1_split_sum(acc);
    if (valid && lane == 0) st_vec(out + b * out_bstride + i, acc);
}
// look-ups of a launch this small are waited for because of the DEPTH of their chains of additions
static bool fb_uses_split(size_t products) { return g1_stage_uses_quads(products / 2, 1); }
void launch_fk20_part2_fold2(const G1A* table, int W, const Fr* c, const Fr* rev, size_t rstride, G1J* out, size_t out_bstride,
                             size_t k, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_LOOKUP, st);
    if (!k || !batch) return;
    if (fb_uses_split(2 * k * batch)) {
        const unsigned g = grid_for(2 * k * batch * FB_SPLIT, G1_BLOCK);
        if (W == 12) k_fk20_part2_fold2_split<12><<<g, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
        else if (W == 10) k_fk20_part2_fold2_split<10><<<g, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
        else if (W == 8) k_fk20_part2_fold2_split<8><<<g, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
        else k_fk20_part2_fold2_split<4><<<g, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
        g_launch_count++;
        return;
    }
    const unsigned grid = grid_for(2 * k * batch, G1_BLOCK);
    if (W == 12) k_fk20_part2_fold2<12><<<grid, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
    else if (W == 10) k_fk20_part2_fold2<10><<<grid, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
    else if (W == 8) k_fk20_part2_fold2<8><<<grid, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
    else k_fk20_part2_fold2<4><<<grid, G1_BLOCK, 0, st>>>(table, c, rev, rstride, out, out_bstride, k, batch);
    g_launch_count++;
}
void launch_g1_mul_fixed_base(const G1A* table, int W, const Fr* k, int k_is_mont, G1J* out, size_t out_bstride, size_t n, size_t batch,
                              cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_LOOKUP, st);
    if (!n || !batch) return;
    if (fb_uses_split(n * batch)) {
        const unsigned g = grid_for(n * batch * FB_SPLIT, G1_BLOCK);
        if (W == 12) k_g1_mul_fixed_base_split<12><<<g, G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
        else if (W == 10) k_g1_mul_fixed_base_split<10><<<g, G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
        else if (W == 8) k_g1_mul_fixed_base_split<8><<<g, G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
        else k_g1_mul_fixed_base_split<4><<<g, G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
        g_launch_count++;
        return;
    }
    if (W == 12) k_g1_mul_fixed_base<12><<<grid_for(n * batch, G1_BLOCK), G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
    else if (W == 10) k_g1_mul_fixed_base<10><<<grid_for(n * batch, G1_BLOCK), G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
    else if (W == 8) k_g1_mul_fixed_base<8><<<grid_for(n * batch, G1_BLOCK), G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
    else k_g1_mul_fixed_base<4><<<grid_for(n * batch, G1_BLOCK), G1_BLOCK, 0, st>>>(table, k, k_is_mont, out, out_bstride, n, batch);
    g_launch_count++;
}
size_t fixed_base_row_entries(int W) { return (size_t)fb_windows(W) * fb_entries(W); }
size_t fixed_base_table_bytes(size_t n, int W) { return n * fixed_base_row_entries(W) * sizeof(G1A); }
size_t fixed_base_tmp_bytes(size_t n, int W) { return n * (size_t)fb_windows(W) * sizeof(G1J); }

// ------------------------------------------------------------------------------ folds / adds
__global__ void __launch_bounds__(128) k_g1_fold(G1J* data, size_t bstride, size_t half, size_t cnt, size_t batch) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t per = cnt - half;
    if (t >= per * batch) return;
    size_t b = t / per, i = t % per;
    G1J* p = data + b * bstride + i;
    G1J x = ld_vec(p), y = ld_vec(p + half), r;
    g1_add_ni(&r, &x, &y);
    st_vec(p, r);
}
void launch_g1_fold(G1J* data, size_t bstride, size_t half, size_t cnt, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_FOLD, st);
    size_t total = (cnt - half) * batch;
    if (!total) return;
    k_g1_fold<<<grid_for(total, 128), 128, 0, st>>>(data, bstride, half, cnt, batch); g_launch_count++;
}

__global__ void __launch_bounds__(128) k_g1_add_arrays(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src,
                                                       size_t src_estride, size_t src_bstride, size_t n, size_t batch) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    size_t b = t / n, i = t % n;
    G1J* p = dst + b * dst_bstride + i * dst_estride;
    G1J x = ld_vec(p), y = ld_vec(src + b * src_bstride + i * src_estride), r;
    g1_add_ni(&r, &x, &y);
    st_vec(p, r);
}
void launch_g1_add_arrays(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src, size_t src_estride,
                          size_t src_bstride, size_t n, size_t batch, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_FOLD, st);
    if (!n || !batch) return;
    k_g1_add_arrays<<<grid_for(n * batch, 128), 128, 0, st>>>(dst, dst_estride, dst_bstride, src, src_estride, src_bstride, n, batch);
    g_launch_count++;
}

// flag |= 1 if any of the n points is off the curve Y^2 = X^3 + 4 Z^6 (infinity, Z == 0, is on it): the aggregated checks feed
// caller-supplied points into MSMs, whose results mean nothing for points outside the group
__global__ void __launch_bounds__(128) k_g1_on_curve(const G1J* __restrict__ pts, size_t n, uint32_t* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G1J p = ld_vec(pts + i);
    if (p.is_inf()) return;
    const Fp z2 = fp_sqr(p.z), z6 = fp_mul(fp_sqr(z2), z2);
    const Fp rhs = fe_add(fp_mul(fp_sqr(p.x), p.x), fp_mul(fp_const_four(), z6));
    if (fp_sqr(p.y) != rhs) atomicOr(flag, 1u);
}
void launch_g1_on_curve(const G1J* pts, size_t n, uint32_t* flag, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n) return;
    k_g1_on_curve<<<grid_for(n, 128), 128, 0, st>>>(pts, n, flag); g_launch_count++;
}

__global__ void __launch_bounds__(128) k_g1_sub_arrays(G1J* dst, const G1J* src, size_t src_stride, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1J x = ld_vec(dst + i), y = ld_vec(src + i * src_stride), r;
    y.y = fe_neg(y.y);
    g1_add_ni(&r, &x, &y);
    st_vec(dst + i, r);
}
void launch_g1_sub_arrays(G1J* dst, const G1J* src, size_t src_stride, size_t n, cudaStream_t st) {
    ProfScope prof_scope(PROF_G1_FOLD, st);
    if (!n) return;
    k_g1_sub_arrays<<<grid_for(n, 128), 128, 0, st>>>(dst, src, src_stride, n);
    g_launch_count++;
}

__global__ void k_g1_copy(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src, size_t src_estride,
                          size_t src_bstride, size_t n, size_t batch, int bitrev, unsigned logn) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * batch) return;
    size_t b = t / n, i = t % n;
    size_t si = bitrev ? bitrev_u32((uint32_t)i, logn) : i;
    st_vec(dst + b * dst_bstride + i * dst_estride, ld_vec(src + b * src_bstride + si * src_estride));
}
void launch_g1_copy(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src, size_t src_estride,
                    size_t src_bstride, size_t n, size_t batch, int bitrev, unsigned logn, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!n || !batch) return;
    k_g1_copy<<<grid_for(n * batch, 256), 256, 0, st>>>(dst, dst_estride, dst_bstride, src, src_estride, src_bstride, n, batch, bitrev, logn);
    g_launch_count++;
}

// dst[(a B + b) C + k] = src[(b A + a) C + k]: swaps the two leading digits of a three-digit index (parts gathered
// rank-major -> natural position order in the sharded merge of FK20 multi)
__global__ void k_g1_swap_digits(G1J* __restrict__ dst, const G1J* __restrict__ src, size_t A, size_t B, size_t C) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A * B * C) return;
    size_t k = t % C, ab = t / C, b = ab % B, a = ab / B;
    st_vec(dst + t, ld_vec(src + (b * A + a) * C + k));
}
void launch_g1_swap_digits(G1J* dst, const G1J* src, size_t A, size_t B, size_t C, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    if (!(A * B * C)) return;
    k_g1_swap_digits<<<grid_for(A * B * C, 256), 256, 0, st>>>(dst, src, A, B, C); g_launch_count++;
}

__global__ void k_fk20_gather_x(const G1J* __restrict__ S, G1J* __restrict__ work, size_t n, size_t l, size_t off0, size_t files) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t k = n / l;
    if (k < 2 || t >= (k - 1) * files) return;
    size_t f = t / (k - 1), i = t % (k - 1), off = off0 + f;
    st_vec(work + f * 2 * k + i, ld_vec(S + (n - l - 1 - off - i * l)));
}
void launch_fk20_gather_x(const G1J* secret_g1, G1J* work, size_t n, size_t l, size_t off0, size_t files, cudaStream_t st) {
    ProfScope prof_scope(PROF_MISC, st);
    size_t k = n / l;
    if (k < 2 || !files) return;
    k_fk20_gather_x<<<grid_for((k - 1) * files, 256), 256, 0, st>>>(secret_g1, work, n, l, off0, files); g_launch_count++;
}

// ------------------------------------------------------------------------------ self test
__device__ uint32_t st_rand(uint64_t& s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (uint32_t)(s >> 32);
}
// The self test is split into small kernels that hand points over through global memory: one
// large kernel holding many 144-byte locals triggered wrong stack-slot sharing in nvcc 12.9
// (a live point overwritten by a later temporary), see DESIGN.md "compiler notes".
__device__ __forceinline__ void st_scalars(size_t i, uint64_t seed, Fp& a, Fp& b, Fr& c, Fr& d) {
    uint64_t s = seed + 0x9E3779B97F4A7C15ULL * (i + 1);
    for (int k = 0; k < 12; k++) { a.l[k] = st_rand(s); b.l[k] = st_rand(s); }
    a.l[11] &= 0x0fffffffu; b.l[11] &= 0x0fffffffu;
    if (i % 5 == 0) for (int k = 0; k < 12; k++) a.l[k] = FpParams::mod(k) - (k == 0 ? 1u : 0u);
    for (int k = 0; k < 8; k++) { c.l[k] = st_rand(s); d.l[k] = st_rand(s); }
    c.l[7] &= 0x3fffffffu; d.l[7] &= 0x3fffffffu;
}
#define ST_CHECK(idx, cond) do { if (!(cond)) { atomicAdd(mismatch, 1ULL); atomicAdd(mismatch + 1 + (idx), 1ULL); } } while (0)
__global__ void k_selftest_field(size_t n, uint64_t seed, unsigned long long* mismatch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp a, b; Fr c, d;
    st_scalars(i, seed, a, b, c, d);
    ST_CHECK(0, fe_mul(a, b) == fe_mul_portable(a, b));
    ST_CHECK(1, fe_sqr(a) == fe_mul_portable(a, a));
    ST_CHECK(2, fe_mul(c, d) == fe_mul_portable(c, d));
    ST_CHECK(3, fe_sub(fe_add(c, d), d) == c);
}
// WHICH: 0 -> k1 G windowed GLV, 1 -> k2 G windowed GLV, 2 -> (k1 + k2) G double-and-add
template <int WHICH>
__global__ void k_selftest_mul(size_t n, uint64_t seed, G1J* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp a, b; Fr c, d;
    st_scalars(i, seed, a, b, c, d);
    Fr k = WHICH == 1 ? d : (WHICH == 2 ? fe_add(c, d) : c);
    G1J g = g1_generator(), r;
    if (WHICH <= 1) g1_mul_var(&r, &g, k.l);
    else r = g1_mul_simple(g, k.l);
    out[(size_t)WHICH * n + i] = r;
}
__global__ void k_selftest_group(size_t n, const G1J* pts, unsigned long long* mismatch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const G1J p1 = pts[i], p2 = pts[n + i], p3 = pts[2 * n + i];
    G1J t, u;
    g1_add_ni(&t, &p1, &p2);
    ST_CHECK(5, g1_equal(t, p3));                         // windowed GLV: k1 G + k2 G == (k1 + k2) G by double-and-add
    g1_add_sub_ni(&t, &u, &p3, &p2);
    ST_CHECK(6, g1_equal(u, p1));                         // butterfly difference
    g1_add_ni(&u, &p3, &p2);
    ST_CHECK(7, g1_equal(t, u));                          // butterfly sum
    g1_dbl_ni(&t, &p1);
    g1_add_ni(&u, &p1, &p1);                              // addition falling into the doubling branch
    ST_CHECK(8, g1_equal(t, u));
    t = g1_neg(p1);
    g1_add_ni(&u, &p1, &t);
    ST_CHECK(9, u.is_inf());                              // P + (-P)
}
// z^2 (x, y) == (beta x, -y) on the generator
__global__ void k_selftest_endo(G1J* out) {
    G1J g = g1_generator();
    uint32_t z2[8] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u, 0, 0, 0, 0};
    out[threadIdx.x] = threadIdx.x == 0 ? g1_endo(g) : g1_mul_simple(g, z2);
}
__global__ void k_selftest_endo_check(const G1J* pts, unsigned long long* mismatch) {
    ST_CHECK(10, g1_equal(pts[0], pts[1]));
}
// host-built scalar programs executed on the device against double-and-add: checks 11 (mode 0), 12 (mode 1)
__global__ void k_selftest_programs(size_t n, const ScalarProgram* progs, const Fr* scalars, unsigned long long* mismatch) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1J g = g1_generator(), got;
    Fr k = scalars[i];
    G1J base = g1_mul_simple(g, k.l);                 // some point other than the generator
    G1J want = g1_mul_simple(base, k.l);
    g1_mul_program(&got, &base, progs + i);
    if (!g1_equal(got, want)) { atomicAdd(mismatch, 1ULL); atomicAdd(mismatch + 1 + 11 + progs[i].mode, 1ULL); }
}
void launch_selftest_programs(size_t n, const ScalarProgram* progs, const Fr* scalars, unsigned long long* d_mismatch, cudaStream_t st) {
    if (!n) return;
    k_selftest_programs<<<grid_for(n, 64), 64, 0, st>>>(n, progs, scalars, d_mismatch); g_launch_count++;
}
void launch_selftest(size_t n, uint64_t seed, unsigned long long* d_mismatch, G1J* d_scratch /* 4 n */, cudaStream_t st) {
    if (!n) return;
    k_selftest_field<<<grid_for(n, 64), 64, 0, st>>>(n, seed, d_mismatch);
    k_selftest_mul<0><<<grid_for(n, 64), 64, 0, st>>>(n, seed, d_scratch);
    k_selftest_mul<1><<<grid_for(n, 64), 64, 0, st>>>(n, seed, d_scratch);
    k_selftest_mul<2><<<grid_for(n, 64), 64, 0, st>>>(n, seed, d_scratch);
    k_selftest_group<<<grid_for(n, 64), 64, 0, st>>>(n, d_scratch, d_mismatch);
    k_selftest_endo<<<1, 2, 0, st>>>(d_scratch);
    k_selftest_endo_check<<<1, 1, 0, st>>>(d_scratch, d_mismatch);
    g_launch_count += 8;
}

__global__ void k_fp_mul_probe(uint32_t* buf, int iters) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    Fp x, y;
    for (int k = 0; k < 12; k++) { x.l[k] = buf[k] + (uint32_t)i; y.l[k] = buf[12 + k] ^ (uint32_t)i; }
    x.l[11] &= 0x0fffffffu; y.l[11] &= 0x0fffffffu;
    for (int k = 0; k < iters; k++) { x = fe_mul(x, y); y = fe_mul(y, x); }
    uint32_t acc = 0;
    for (int k = 0; k < 12; k++) acc ^= x.l[k] ^ y.l[k];
    if (acc == 0x12345678u) buf[24] = acc;   // keep the chain alive
}
void launch_fp_mul_probe(uint32_t* d_buf, size_t threads, int iters, cudaStream_t st) {
    k_fp_mul_probe<<<grid_for(threads, 128), 128, 0, st>>>(d_buf, iters); g_launch_count++;
}

}  // namespace b200
