// kernels.h -- launcher interface between the C-ABI translation unit (api.cu) and the kernel
// translation units (kernels_fr.cu, kernels_g1.cu).  Internal; the public boundary is
// include/b200_kzg.h.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "g1.cuh"

namespace b200 {

// every launcher bumps this (gpu_launches in bench.py): kernels launched by the calling thread, so that a
// before/after difference around a call is exact whatever other threads do
extern thread_local unsigned long long g_launch_count;

// Optional per-kernel-class device timing (bench.py's roofline): when enabled, every launcher
// brackets its launches with CUDA events on the launching stream.
enum ProfCat { PROF_FR_NTT = 0, PROF_G1_FFT_STAGE, PROF_G1_MUL, PROF_G1_FOLD, PROF_MISC, PROF_G1_LOOKUP, PROF_G1_MSM, PROF_NCAT };
extern std::atomic<bool> g_prof_on;
long prof_begin_event(int cat, cudaStream_t st);   // returns the record's index (-1: not recorded)
void prof_end_event(long rec, cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    long rec;
    ProfScope(int cat, cudaStream_t s) : st(s), rec(g_prof_on.load(std::memory_order_relaxed) ? prof_begin_event(cat, s) : -1) {}
    ~ProfScope() { if (rec >= 0) prof_end_event(rec, st); }
};

// ---------------------------------------------------------------- Fr
// Device-side FFTSettings tables (fft.go:34-42), Montgomery form.
struct FrDomain {
    unsigned max_scale = 0;
    uint64_t max_width = 0;
    Fr* expanded = nullptr;   // [max_width + 1]  w^i
    Fr* reverse = nullptr;    // [max_width + 1]  w^-i
    // compact per-scale half tables: scale s (n = 2^s) lives at offset n/2, length n/2:
    // tw_fwd[n/2 + j] = w_n^j, tw_inv[n/2 + j] = w_n^-j  (j < n/2); contiguous so one bulk copy
    // stages a transform's twiddles in shared memory.
    Fr* tw_fwd = nullptr;     // [max_width]
    Fr* tw_inv = nullptr;     // [max_width]
};

void launch_fr_to_mont(const uint64_t* in_canon, Fr* out, size_t n, cudaStream_t st);
void launch_fr_from_mont(const Fr* in, uint64_t* out_canon, size_t n, cudaStream_t st);
// out[b][i] = scale * NTT(in[b][:])[i]; natural order in and out; n = 2^logn <= max_width; tmp
// holds batch * n elements (used when the transform takes two passes); in may equal out.
// `scale` (Montgomery) is multiplied into every output (n^-1 for the inverse transform).
void launch_fr_ntt(const FrDomain& dom, const Fr* in, Fr* out, Fr* tmp, unsigned logn, size_t batch, bool inverse,
                   const Fr* scale_or_null, cudaStream_t st);
// das_extension.go:71-84 in place on Montgomery values [batch][2^logn]; inv_n = (2^logn)^-1 (Montgomery)
void launch_das_fft_extension(const FrDomain& dom, Fr* vals, unsigned logn, size_t batch, const Fr& inv_n, cudaStream_t st);
// zero_poly.go:116-217 (evaluation side): zero_eval[b][j] = prod_i (w^j - w^missing[b][i]) for j < n;
// missing: [batch][miss_pitch] u32 indices, nmiss[b] of them valid; partial: [batch][segments][n] scratch
size_t zero_eval_segments(size_t max_missing);
// product-tree route for large missing sets: zero_poly[b][0..n) (Montgomery coefficients); mp = zero_poly_tree_size(max
// missing) <= n padded roots per list; coef_a, coef_b: batch * mp scratch each, padded, ntt_tmp: batch * 2 mp each
size_t zero_poly_tree_size(size_t max_missing);
void launch_missing_lists(const uint8_t* present, size_t n, size_t batch, uint32_t* missing, size_t pitch, uint32_t* nmiss, cudaStream_t st);
void launch_zero_poly_tree(const FrDomain& dom, size_t n, size_t batch, const uint32_t* d_missing, const uint32_t* d_nmiss,
                           size_t miss_pitch, size_t mp, Fr* coef_a, Fr* coef_b, Fr* padded, Fr* ntt_tmp, Fr* zero_poly,
                           cudaStream_t st);
void launch_zero_eval(const FrDomain& dom, size_t n, size_t batch, const uint32_t* d_missing, const uint32_t* d_nmiss,
                      size_t miss_pitch, size_t max_missing, Fr* partial, Fr* zero_eval, cudaStream_t st);
void launch_fr_mul_masked(Fr* dst, const Fr* a, const Fr* c, const uint8_t* present, size_t total, cudaStream_t st);
void launch_fr_mul_table(Fr* v, const Fr* table, size_t n, size_t batch, cudaStream_t st);   // v[b][i] *= table[i]
void launch_fr_div(Fr* a, const Fr* c, size_t total, cudaStream_t st);                        // a[i] /= c[i]  (x / 0 = 0)
void launch_recover_check(const Fr* rec, const Fr* samples, const Fr* zero_eval, const uint8_t* present, size_t n, size_t batch,
                          uint32_t* flags, cudaStream_t st);
// fk20_single.go:106-119 toeplitzCoeffsStep over a batch: poly[b][n] canonical -> out[b][2n] Montgomery
void launch_toeplitz_coeffs(const uint64_t* polys_canon, Fr* out, size_t n, size_t batch, cudaStream_t st);
// fk20_single.go:89-103 toeplitzCoeffsStepStrided for every offset: out[b][off][2k] Montgomery
void launch_toeplitz_coeffs_strided(const uint64_t* polys_canon, Fr* out, size_t n, size_t chunk_len, size_t batch,
                                    cudaStream_t st);
// out[i] = s^i (canonical limbs) for i < n; d_sq[j] = s^(2^j) (Montgomery), 40 entries
void launch_fr_powers(const Fr* d_sq, Fr* out_canon, size_t n, cudaStream_t st);
// bls/bignum_all.go:12-35 ValidFr over batch x n canonical elements: ok[b] (pre-set to 1) is cleared when
// an element of blob b is >= r  (eth/helpers.go:264-273 BlobToPolynomial)
void launch_fr_check_canonical(const uint64_t* vals, size_t n, size_t batch, uint32_t* ok, cudaStream_t st);
// eth/helpers.go:179-203 ComputeKZGProof, field side: y[b] = f_b(z_b) (bls/globals.go:106-153) and the quotient
// q[b][i] = (f[b][i] - y[b]) / (D[i] - z[b]) on the domain of size 2^logn in reverse bit order (bitrev != 0: eth DomainFr) or
// natural order; ok[b] (may be null) cleared if z[b] is in the domain; q_canon == null: evaluation only
void launch_eval_form_quotient(const FrDomain& dom, const uint64_t* f_canon, const uint64_t* z_canon, unsigned logn, size_t batch, int bitrev,
                               const Fr& inv_n, Fr* inv_den, Fr* partial, Fr* y_mont, uint64_t* y_canon_or_null, uint64_t* q_canon,
                               uint32_t* ok, cudaStream_t st);
// kzg_multi_proofs.go:57-71: coeffs[b][i] (Montgomery in, canonical out) /= x[b]^i; x_pow_n[b] = x[b]^(2^logn) (canonical)
void launch_unscale_coset(Fr* coeffs, const uint64_t* x_canon, unsigned logn, size_t batch, uint64_t* x_pow_n_canon, cudaStream_t st);
// pointwise helpers on Montgomery arrays
void launch_fr_mul_arrays(Fr* dst, const Fr* a, const Fr* b, size_t n, cudaStream_t st);   // dst = a * b
void launch_fr_mul_even_odd(Fr* v, const Fr& even, const Fr& odd, size_t total, cudaStream_t st);   // v[i] *= i even ? even : odd

// ---------------------------------------------------------------- G1
void launch_g1_from_abi(const uint64_t* in, G1J* out, size_t n, cudaStream_t st);
// out[b * n + i] = in[b * bstride + idx(i) * estride], idx = bit reversal over logn bits if bitrev
void launch_g1_to_abi(const G1J* in, uint64_t* out, size_t n, size_t batch, size_t estride, size_t bstride, int bitrev,
                      unsigned logn, cudaStream_t st);
void launch_g1_fill_infinity(G1J* p, size_t n, cudaStream_t st);
// same addressing as launch_g1_to_abi, output = 48-byte compressed points (bls/bls_kilic.go:114 ToCompressedG1)
void launch_g1_compress(const G1J* in, uint8_t* out48, size_t n, size_t batch, size_t estride, size_t bstride, int bitrev,
                        unsigned logn, cudaStream_t st);
// one radix-2 stage over batch transforms of 2 * n_half points each; element i of blob b lives at
// data[b * bstride + i * estride]; m = half block length of this stage.  DIT: (x0 + w x1, x0 - w x1),
// DIF: (x0 + x1, w (x0 - x1)), w = progs[j * prog_stride] for position j inside the block.
// across_blocks != 0 (needs n_half / m a multiple of 32): lanes run over the blocks of the stage, so that the 32 lanes of a warp
// share the twiddle w^j even for a single transform; progs must then be the sparse (mode 1) programs
void launch_g1_fft_stage(G1J* data, size_t n_half, size_t batch, size_t m, size_t estride, size_t bstride, bool dif,
                         const ScalarProgram* progs, size_t prog_stride, cudaStream_t st, int across_blocks = 0, size_t jmul = 1, size_t bmul = 0,
                         size_t joff = 0);   // twiddle index (j jmul + b bmul + joff) prog_stride, b = index in the batch
// whether launch_g1_fft_stage runs one butterfly per quad of lanes for this shape (then across_blocks needs n_half / m to be a
// multiple of 8 only, instead of 32)
bool g1_stage_uses_quads(size_t n_half, size_t batch);
void g1_set_quad_enabled(bool enabled);   // process wide (b200_set_latency_mode)
void g1_set_quad_allowed(bool allowed);   // per calling thread; the API layer clears it when several calls are in flight
// dst[(a B + b) C + k] = src[(b A + a) C + k]
void launch_g1_swap_digits(G1J* dst, const G1J* src, size_t A, size_t B, size_t C, cudaStream_t st);
// DIF stage of one transform keeping one half of the outputs: out[i] = in[i] + in[i + m] (lower = 0) or
// progs[i * prog_stride] * (in[i] - in[i + m]) (lower = 1), i < m
void launch_g1_dif_half_stage(const G1J* in, G1J* out, size_t m, int lower, const ScalarProgram* progs, size_t prog_stride, cudaStream_t st);
// out[b * out_bstride + i] = k[b * n + i] * pts[b * pts_bstride + i]   (pts_bstride = 0: shared bases)
void launch_g1_mul_var(const G1J* pts, size_t pts_bstride, const Fr* k, int k_is_mont, G1J* out, size_t out_bstride,
                       size_t n, size_t batch, cudaStream_t st);
// fixed-base window tables (signed W-bit windows, W = 8 or 4, affine entries): build and use
size_t fixed_base_row_entries(int W);                 // table entries per base
size_t fixed_base_table_bytes(size_t n, int W);
size_t fixed_base_tmp_bytes(size_t n, int W);
void launch_fixed_base_table(const G1J* pts, size_t n, G1J* bases_tmp, G1A* table, int W, cudaStream_t st);
void launch_g1_mul_fixed_base(const G1A* table, int W, const Fr* k, int k_is_mont, G1J* out, size_t out_bstride, size_t n, size_t batch,
                              cudaStream_t st);
// FK20Single: out[b][i] = c[b][i] X[i] for even i; the odd slots receive the result of the first two inverse DIF
// stages over x[m] = c[2m+1] X[2m+1] (m < k, k >= 8), computed as 4-term fixed-base sums (see kernels_g1.cu)
void launch_fk20_part2_fold2(const G1A* table, int W, const Fr* c_mont, const Fr* rev, size_t rstride, G1J* out, size_t out_bstride,
                             size_t k, size_t batch, cudaStream_t st);
// out[b * bstride + i * estride] = progs[idx(i) * prog_stride] * same element (in place)
void launch_g1_mul_programs(G1J* data, size_t n, size_t batch, size_t estride, size_t bstride,
                            const ScalarProgram* progs, size_t prog_stride, int bitrev, unsigned logn, cudaStream_t st);
// data[b * bstride + i] += data[b * bstride + i + half]  for i < cnt - half
void launch_g1_fold(G1J* data, size_t bstride, size_t half, size_t cnt, size_t batch, cudaStream_t st);
// dst[b*dst_bstride + i*dst_estride] += src[b*src_bstride + i*src_estride]
void launch_g1_add_arrays(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src, size_t src_estride,
                          size_t src_bstride, size_t n, size_t batch, cudaStream_t st);
// dst[i] -= src[i * src_stride]   (contiguous dst; src_stride in points)
void launch_g1_sub_arrays(G1J* dst, const G1J* src, size_t src_stride, size_t n, cudaStream_t st);
void launch_g1_on_curve(const G1J* pts, size_t n, uint32_t* flag, cudaStream_t st);   // flag |= 1 if a point is off the curve
// bls/bls_kilic.go:118-121 FromCompressedG1 over an array: flags, x < p, curve equation, prime-order subgroup;
// out[i] = ABI point (canonical, Z = 1; all zero for infinity and for rejected encodings), status[i] = 0 ok / 1 / 2 / 3
void launch_g1_decompress(const uint8_t* in48, uint64_t* out_abi, uint32_t* status, size_t n, cudaStream_t st);
// dst[b*dst_bstride + i*dst_estride] = src[b*src_bstride + idx(i)*src_estride]
void launch_g1_copy(G1J* dst, size_t dst_estride, size_t dst_bstride, const G1J* src, size_t src_estride,
                    size_t src_bstride, size_t n, size_t batch, int bitrev, unsigned logn, cudaStream_t st);
// kzg.go:57-61 / 103-109 for the chunk offsets off0 .. off0 + files - 1: work[f * 2k + i] = S[n - l - 1 - (off0 + f) - i l]
// for i < k - 1 (k = n / l); the rest of work (pre-filled) stays infinity
void launch_fk20_gather_x(const G1J* secret_g1, G1J* work, size_t n, size_t l, size_t off0, size_t files, cudaStream_t st);
// Pippenger bucket MSM over variable bases (kernels_msm.cu): *out = sum_i k[i] pts[i]; workspace of msm_workspace_bytes(n)
size_t msm_workspace_bytes(size_t n);
void launch_g1_msm(const G1J* pts, const Fr* k, int k_is_mont, size_t n, void* workspace, G1J* out, cudaStream_t st);
// self test: device field + group law against portable forms; returns mismatches
void launch_selftest(size_t n, uint64_t seed, unsigned long long* d_mismatch, G1J* d_scratch /* 4 n points */, cudaStream_t st);
void launch_selftest_programs(size_t n, const ScalarProgram* progs, const Fr* scalars_canon, unsigned long long* d_mismatch, cudaStream_t st);
// throughput probe: `iters` dependent Fp multiplications per thread
void launch_fp_mul_probe(uint32_t* d_buf, size_t threads, int iters, cudaStream_t st);

}  // namespace b200
