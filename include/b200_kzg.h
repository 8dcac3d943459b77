/*
 * b200_kzg.h -- C ABI of the B200-native KZG / FFT engine (libb200kzg.so).
 *
 * This is the drop-in boundary for protolambda/go-kzg's hot path: a new `bls/` build tag
 * (`bignum_b200`) binds these symbols over cgo and keeps the Go
 * FFTSettings / KZGSettings / FK20SingleSettings / FK20MultiSettings API unchanged.
 * Every entry point cites the reference interface (file:line under /root/reference) it
 * replaces.  No torch types, plain pointers and sizes only.
 *
 * Encodings (ours to choose: bls.Fr / bls.G1Point are opaque outside package bls):
 *   Fr       4 x uint64 little-endian limbs, canonical residue in [0, r)   (== FrTo32 bytes)
 *   G1       Jacobian X, Y, Z, each 6 x uint64 little-endian canonical limbs in [0, p);
 *            point at infinity <=> Z == 0, so the all-zero value is infinity (Go zero value).
 *   G1 compressed: 48 bytes, ZCash/IETF form (bls/bls_kilic.go:114-121).
 * Arrays are contiguous ([]bls.Fr / []bls.G1Point pass as &s[0] without copying).
 *
 * Ownership: the caller allocates all inputs and outputs; the library never keeps a caller
 * pointer after return.  Device tables live in the opaque handles.  Handles are immutable
 * after creation (lazily built tables are published under a lock and never freed or moved while the handle
 * lives) and may be used from any thread concurrently; each call sets its CUDA device itself and runs on a
 * pooled non-blocking stream of its own, so concurrent callers do not serialise on the default stream.
 *
 * Errors: every function returns a b200_status.  The Go shim maps them back to the
 * reference's error/panic split.  There is NO CPU fallback: without a
 * usable CUDA device every compute entry point returns B200_ERR_CUDA / B200_ERR_NO_DEVICE.
 */
#ifndef B200_KZG_H
#define B200_KZG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B200_OK = 0,
    B200_ERR_TOO_LARGE = 1,     /* "got %d values but only have %d roots of unity" fft_fr.go:57-59,78-80; fft_g1.go:60-62 */
    B200_ERR_NOT_POW2 = 2,      /* "not a power of two" fft_fr.go:81-83; fft_g1.go:63-65; kzg.go:47,77 */
    B200_ERR_LEN_MISMATCH = 3,  /* kzg.go:22-24; fk20_single.go:60-62; bls/bls_kilic.go:133-135 */
    B200_ERR_BAD_INPUT = 4,     /* e.g. "second half should be zeroed" fk20_single.go:150-154; bad point encoding */
    B200_ERR_CUDA = 5,          /* CUDA runtime failure (see b200_last_cuda_error) */
    B200_ERR_NO_DEVICE = 6,     /* no CUDA device / library built without one */
    B200_ERR_TOO_SMALL = 7,     /* kzg.go:25-27,50-52; das_extension.go:72-74 */
    B200_ERR_RECOVERY = 8,      /* recover_from_samples.go:103-107 reconstructed data mismatch */
    B200_ERR_ZERO_EVAL = 9      /* recover_from_samples.go:54-58 "bad zero eval" */
} b200_status;

const char* b200_strerror(int status);
const char* b200_last_cuda_error(void);
int b200_device_count(void);
/* Latency mode of small G1 transforms (one polynomial per call): 1 (default) = automatic -- a call that finds at most one other
 * call in flight spends four lanes per butterfly (FK20Single of one n = 4096 polynomial: 37 -> 23 ms); 0 = never (servers
 * that always run many callers and want the aggregate rate).  Results are identical either way. */
int b200_set_latency_mode(int mode);
/* Selects the CUDA device used by handles created afterwards on this thread (default 0). */
int b200_set_device(int device);

typedef struct b200_fs b200_fs; /* FFTSettings          fft.go:34-42   */
typedef struct b200_ks b200_ks; /* KZGSettings          kzg.go:11-19   */
typedef struct b200_fk b200_fk; /* FK20Single/MultiSettings kzg.go:38-41,66-71 */

/* ------------------------------------------------------------------ level 1: package bls ---
 * Host-side scalar operations so that the Go facade compiles and behaves (API completeness;
 * none of these is on the hot path -- the hot path crosses cgo at batch granularity below). */
void b200_fr_add(uint64_t dst[4], const uint64_t a[4], const uint64_t b[4]); /* bls/bignum_kilic.go:99  AddModFr */
void b200_fr_sub(uint64_t dst[4], const uint64_t a[4], const uint64_t b[4]); /* bls/bignum_kilic.go:95  SubModFr */
void b200_fr_mul(uint64_t dst[4], const uint64_t a[4], const uint64_t b[4]); /* bls/bignum_kilic.go:109 MulModFr */
void b200_fr_inv(uint64_t dst[4], const uint64_t a[4]);                      /* bls/bignum_kilic.go:113 InvModFr */
void b200_fr_div(uint64_t dst[4], const uint64_t a[4], const uint64_t b[4]); /* bls/bignum_kilic.go:103 DivModFr */
void b200_fr_batch_inv(uint64_t* vals, size_t n);                            /* bls/bignum_kilic.go:117 BatchInvModFr (in place) */
int b200_fr_valid(const uint8_t le32[32]);                                   /* bls/bignum_all.go:12-35 ValidFr */
void b200_fr_root_of_unity(unsigned scale, uint64_t out[4]);                 /* bls/globals.go:27-60 Scale2RootOfUnity */
void b200_g1_generator(uint64_t out[18]);                                    /* bls/bls_kilic.go:23 GenG1 */
void b200_g1_add(uint64_t dst[18], const uint64_t a[18], const uint64_t b[18]);     /* bls/bls_kilic.go:47 AddG1 */
void b200_g1_sub(uint64_t dst[18], const uint64_t a[18], const uint64_t b[18]);     /* bls/bls_kilic.go:51 SubG1 */
void b200_g1_neg(uint64_t dst[18]);                                                 /* bls/bls_kilic.go:63 NegG1 (in place) */
void b200_g1_mul(uint64_t dst[18], const uint64_t a[18], const uint64_t k[4]);      /* bls/bls_kilic.go:41 MulG1 */
int b200_g1_equal(const uint64_t a[18], const uint64_t b[18]);                      /* bls/bls_kilic.go:106 EqualG1 */
void b200_g1_to_compressed(uint8_t out[48], const uint64_t p[18]);                  /* bls/bls_kilic.go:114 ToCompressedG1 */
int b200_g1_from_compressed(uint64_t out[18], const uint8_t in[48]);                /* bls/bls_kilic.go:118 FromCompressedG1 */
void b200_g1_to_compressed_many(uint8_t* out, const uint64_t* pts, size_t n);
int b200_g1_from_compressed_many(uint64_t* out, const uint8_t* in, size_t n);

/* G2 and the pairing: HOST code (go_kzg_b200/csrc/pairing.h), verification side only -- two G2 operations and one pairing
 * check per proof.  G2 = Jacobian X, Y, Z over Fp2, each coordinate (c0, c1) of 6 x uint64 canonical limbs: 36 x uint64 =
 * 288 bytes, infinity <=> Z == 0 (all-zero value = infinity).  Compressed: 96 bytes, ZCash form (x.c1 || x.c0). */
void b200_g2_generator(uint64_t out[36]);                                           /* bls/bls_kilic.go:24 GenG2 */
void b200_g2_add(uint64_t dst[36], const uint64_t a[36], const uint64_t b[36]);     /* bls/bls_kilic.go:86 AddG2 */
void b200_g2_sub(uint64_t dst[36], const uint64_t a[36], const uint64_t b[36]);     /* bls/bls_kilic.go:90 SubG2 */
void b200_g2_neg(uint64_t dst[36]);                                                 /* bls/bls_kilic.go:94 NegG2 (in place) */
void b200_g2_mul(uint64_t dst[36], const uint64_t a[36], const uint64_t k[4]);      /* bls/bls_kilic.go:80 MulG2 */
int b200_g2_equal(const uint64_t a[36], const uint64_t b[36]);                      /* bls/bls_kilic.go:110 EqualG2 */
void b200_g2_to_compressed(uint8_t out[96], const uint64_t p[36]);                  /* bls/bls_kilic.go:123 ToCompressedG2 */
/* bls/bls_kilic.go:127 FromCompressedG2: flags, coordinates < p, curve equation, prime-order subgroup; else B200_ERR_BAD_INPUT */
int b200_g2_from_compressed(uint64_t out[36], const uint8_t in[96]);
/* Affine coordinates, canonical limbs (G1: x, y; G2: x.c0, x.c1, y.c0, y.c1; all zero for infinity): what StrG1 / StrG2 print
 * (bls/bls_kilic.go:55-61, 96-102). */
void b200_g1_to_affine(const uint64_t p[18], uint64_t xy[12]);
void b200_g2_to_affine(const uint64_t p[36], uint64_t xy[24]);
/* setup.go:9-26 GenerateTestingSetup, G2 half: out[i] = secret^i * GenG2 (host, one scalar multiplication per entry) */
int b200_generate_testing_setup_g2(const uint64_t secret[4], size_t n, uint64_t* out);
/* bls/bls_kilic.go:152-158 PairingsVerify: *ok = (e(a1, a2) == e(b1, b2)).  Non-canonical coordinates or points off their
 * curve -> B200_ERR_BAD_INPUT. */
int b200_pairings_verify(const uint64_t a1[18], const uint64_t a2[36], const uint64_t b1[18], const uint64_t b2[36], int* ok);
/* e(p, q) itself: the 12 canonical Fp coefficients a_0, b_0, .., a_5, b_5 (6 x uint64 each) of sum (a_i + b_i u) w^i in
 * Fp2[w] / (w^6 - (1 + u)), Fp2 = Fp[u] / (u^2 + 1); exponent (p^12 - 1) / r exactly (tests compare it with an independent
 * restatement). */
int b200_pairing(const uint64_t p[18], const uint64_t q[36], uint64_t out[72]);

/* bls/bls_kilic.go:118-121 FromCompressedG1 over an array ON THE DEVICE (eth/globals.go:33-49 decodes 3 x 4096 points
 * at start-up; larger setups hold millions): flags, x < p, curve equation, prime-order subgroup.  ok (may be NULL)
 * gets 1 per accepted point; a rejected encoding -> B200_ERR_BAD_INPUT with that output zeroed. */
int b200_g1_from_compressed_batch(const uint8_t* in48, size_t n, uint64_t* out, uint8_t* ok);

/* Device MSM.  bls/bls_kilic.go:132-150 LinCombG1: sum_i scalars[i] * points[i]; n == 0 gives
 * infinity (bls/bls_test.go:69-77).  Pippenger bucket method (kernels_msm.cu): GLV halves, signed windows,
 * per-window buckets accumulated in shared memory, warp-shuffle running-sum reduction; below 32 terms the
 * per-term windowed multiplication + fold tree is used instead. */
int b200_g1_lincomb(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t out[18]);
/* Device batch MulG1: out[i] = scalars[i] * points[i] (the loop at fk20_single.go:72-74). */
int b200_g1_mul_many(const uint64_t* points, const uint64_t* scalars, size_t n, uint64_t* out);

/* ------------------------------------------------------------------ FFTSettings ------------ */
int b200_fft_settings_new(uint8_t max_scale, b200_fs** out);  /* fft.go:44-61 NewFFTSettings */
void b200_fft_settings_free(b200_fs* fs);
uint64_t b200_fs_max_width(const b200_fs* fs);
/* ExpandedRootsOfUnity (reverse == 0) or ReverseRootsOfUnity: max_width + 1 entries (fft.go:21-32) */
int b200_fs_roots(const b200_fs* fs, int reverse, uint64_t* out);

/* fft_fr.go:55-74 FFT: n <= MaxWidth values, zero-padded to the next power of two;
 * out holds next_pow2(n) entries, natural order in and out. */
int b200_fft_fr(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out);
/* fft_fr.go:76-105 InplaceFFT: no padding; n not a power of two -> B200_ERR_NOT_POW2 (:81-83); n == 0 ->
 * B200_ERR_BAD_INPUT (the reference divides by n: run-time panic).  vals may equal out. */
int b200_inplace_fft_fr(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out);
/* `batch` independent transforms of identical length n (contiguous). */
int b200_fft_fr_batch(b200_fs* fs, const uint64_t* vals, size_t n, size_t batch, int inverse, uint64_t* out);
/* fft_g1.go:58-94 FFTG1: n must be a power of two <= MaxWidth. */
int b200_fft_g1(b200_fs* fs, const uint64_t* vals, size_t n, int inverse, uint64_t* out);
int b200_fft_g1_batch(b200_fs* fs, const uint64_t* vals, size_t n, size_t batch, int inverse, uint64_t* out);
/* fk20_single.go:59-77 ToeplitzPart2 as a stand-alone method: hExtFFT[i] = FFT(toeplitzCoeffs)[i] * xExtFFT[i] for
 * caller-held points; n_coeffs != n_points -> B200_ERR_LEN_MISMATCH (panic :60-62).  The FK20 entry points below run
 * a fused form over the settings' resident xExtFFT instead. */
int b200_toeplitz_part2(b200_fs* fs, const uint64_t* toeplitz_coeffs, size_t n_coeffs, const uint64_t* x_ext_fft, size_t n_points,
                        uint64_t* h_ext_fft);
/* fk20_single.go:80-87 ToeplitzPart3: FFTG1(hExtFFT, inverse)[:n2 / 2]; out holds n2 / 2 points. */
int b200_toeplitz_part3(b200_fs* fs, const uint64_t* h_ext_fft, size_t n2, uint64_t* out);
/* bls/globals.go:106-153 EvaluatePolyInEvaluationForm for `batch` polynomials of n = 2^k evaluations on the settings'
 * order-n domain, natural order (rootsOfUnity = ExpandedRootsOfUnity, scale = log2(MaxWidth / n)) or reverse bit order
 * (eth DomainFr, eth/globals.go:60-67); xs[b] = evaluation point, y[b] = result (canonical). */
int b200_evaluate_poly_in_evaluation_form_batch(b200_fs* fs, const uint64_t* polys, const uint64_t* xs, size_t n, size_t batch,
                                                int reverse_bit_order, uint64_t* y);
/* das_extension.go:71-84 DASFFTExtension: in place; requires 2 n <= MaxWidth (and, as in the
 * reference, is only meaningful for MaxWidth == 2 n). */
int b200_das_fft_extension(b200_fs* fs, uint64_t* vals, size_t n);
int b200_das_fft_extension_batch(b200_fs* fs, uint64_t* vals, size_t n, size_t batch);
/* das_extension.go:71-84 over G1Points -- the "G1 version of the DAS extension FFT" the reference leaves as a TODO at
 * fk20_multi.go:96: odd-index evaluations from the n even-index ones, in place; same checks as the Fr form. */
int b200_das_fft_extension_g1(b200_fs* fs, uint64_t* vals, size_t n);
/* zero_poly.go:116-217 ZeroPolyViaMultiplication -> (zeroEval[length], zeroPoly[length]) */
int b200_zero_poly_via_multiplication(b200_fs* fs, const uint64_t* missing_indices, size_t n_missing, size_t length,
                                      uint64_t* zero_eval, uint64_t* zero_poly);
/* recover_from_samples.go:42-109 RecoverPolyFromSamples with ZeroPolyViaMultiplication;
 * samples is []*bls.Fr flattened by the shim: present[i] == 0 <=> samples[i] == nil. */
int b200_recover_poly_from_samples(b200_fs* fs, const uint64_t* samples, const uint8_t* present, size_t n, uint64_t* out);
int b200_recover_poly_from_samples_batch(b200_fs* fs, const uint64_t* samples, const uint8_t* present, size_t n,
                                         size_t batch, uint64_t* out);

/* ------------------------------------------------------------------ KZGSettings ------------ */
/* kzg.go:21-36 NewKZGSettings: n_g1 != n_g2 -> LEN_MISMATCH; n_g1 < MaxWidth -> TOO_SMALL.
 * Only the length of the G2 half is checked here; b200_kzg_settings_set_secret_g2 hands the points over (verification only). */
int b200_kzg_settings_new(b200_fs* fs, const uint64_t* secret_g1, size_t n_g1, size_t n_g2, b200_ks** out);
void b200_kzg_settings_free(b200_ks* ks);
/* kzg_single_proofs.go:17-19 CommitToPoly = LinCombG1(SecretG1[:n], coeffs) */
int b200_commit_to_poly(b200_ks* ks, const uint64_t* coeffs, size_t n, uint64_t out[18]);
int b200_commit_to_poly_batch(b200_ks* ks, const uint64_t* coeffs, size_t n, size_t batch, uint64_t* out);

/* ------------------------------------------------------------------ verification, G1 side ---
 * kzg_single_proofs.go:57-75 CheckProofSingle: out[i] = commitment[i] - y[i] G.  The G2 arithmetic and the pairing
 * stay with the caller's backend (bls.PairingsVerify): e(out[i], [1]_2) == e(proof[i], [s - x[i]]_2). */
int b200_check_proof_single_g1_batch(const uint64_t* commitments, const uint64_t* ys, size_t batch, uint64_t* out);
/* kzg_multi_proofs.go:47-88 CheckProofMulti for `batch` samples of n = len(ys) values each (power of two):
 * out_g1[b] = commitment[b] - [interpolation_polynomial_b(s)]_1, x_pow_n[b] = x[b]^n; the caller checks
 * e(out_g1[b], [1]_2) == e(proof[b], SecretG2[n] - [x^n]_2).  n > MaxWidth -> TOO_LARGE (the reference panics). */
int b200_check_proof_multi_g1_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* xs, const uint64_t* ys, size_t n,
                                    size_t batch, uint64_t* out_g1, uint64_t* x_pow_n);

/* KZGSettings.SecretG2 (kzg.go:14-16): host copy of the n first points, read only by the two checks below (SecretG2[1],
 * SecretG2[len(ys)]).  Call once, before the handle is shared between threads.  A point is validated when a check reads it
 * (non-canonical coordinates or off the twist -> B200_ERR_BAD_INPUT from that check). */
int b200_kzg_settings_set_secret_g2(b200_ks* ks, const uint64_t* secret_g2, size_t n);
/* kzg_single_proofs.go:57-75 CheckProofSingle, complete: G1 side on the device, [s - x]_2 and the pairing check on the host.
 * *ok / ok[i] = 1 if the proof verifies.  SecretG2 not set (fewer than 2 points) -> B200_ERR_TOO_SMALL. */
int b200_check_proof_single(b200_ks* ks, const uint64_t commitment[18], const uint64_t proof[18], const uint64_t x[4],
                            const uint64_t y[4], int* ok);
int b200_check_proof_single_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                  const uint64_t* ys, size_t batch, uint8_t* ok);
/* kzg_multi_proofs.go:47-88 CheckProofMulti, complete (n = len(ys), power of two; needs SecretG2[n]); the batch form takes
 * `batch` samples (commitment, proof, x, ys[n]) and spreads the pairing checks over the host cores. */
int b200_check_proof_multi(b200_ks* ks, const uint64_t commitment[18], const uint64_t proof[18], const uint64_t x[4],
                           const uint64_t* ys, size_t n, int* ok);
int b200_check_proof_multi_batch(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                 const uint64_t* ys, size_t n, size_t batch, uint8_t* ok);

/* Aggregated forms (not in the reference): all `batch` proofs with ONE pairing.  With caller-supplied unpredictable non-zero
 * scalars rs[i] (canonical Fr, from a CSPRNG), the per-proof equations are folded into
 *   e(sum r_i A_i + sum r_i c_i proof_i, [1]_2) == e(sum r_i proof_i, [t]_2)   (A_i, c_i, t as in the functions above):
 * three device MSMs of `batch` terms and one host pairing.  *ok = 1 iff the folded equation holds (a forged proof passes with
 * probability ~ batch / r).  Use the per-proof forms to find WHICH proof of a failing batch is bad. */
int b200_check_proof_single_aggregate(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                      const uint64_t* ys, const uint64_t* rs, size_t batch, int* ok);
int b200_check_proof_multi_aggregate(b200_ks* ks, const uint64_t* commitments, const uint64_t* proofs, const uint64_t* xs,
                                     const uint64_t* ys, size_t n, const uint64_t* rs, size_t batch, int* ok);

/* ------------------------------------------------------------------ FK20 ------------------- */
/* kzg.go:43-64 NewFK20SingleSettings(ks, n2) */
int b200_fk20_single_settings_new(b200_ks* ks, size_t n2, b200_fk** out);
/* kzg.go:73-116 NewFK20MultiSettings(ks, n2, chunkLen) */
int b200_fk20_multi_settings_new(b200_ks* ks, size_t n2, size_t chunk_len, b200_fk** out);
/* The same for one rank of the offset-sharded FK20 multi (multi-GPU building blocks below): only the xExtFFT files of
 * the chunk offsets [off_begin, off_end) and their window tables are built and kept; such a handle serves
 * b200_fk20_multi_partial_dev for those offsets and the finish calls, not the whole-polynomial entry points. */
int b200_fk20_multi_settings_new_sharded(b200_ks* ks, size_t n2, size_t chunk_len, size_t off_begin, size_t off_end, b200_fk** out);
/* Setup-time precompute as a cacheable artefact (kzg.go:101-114 costs 16 G1 transforms of 2^17 points
 * per process at config 5): settings from xExtFFT files exported earlier with b200_fk20_x_ext_fft (offsets
 * [off_begin, off_end), (off_end - off_begin) x n2 / chunk_len points, concatenated).  chunk_len = 1 gives single settings. */
int b200_fk20_settings_new_from_x_ext_fft(b200_ks* ks, size_t n2, size_t chunk_len, size_t off_begin, size_t off_end,
                                          const uint64_t* x_ext_fft, b200_fk** out);
void b200_fk20_settings_free(b200_fk* fk);
/* copy of xExtFFT (file `file`, n2 / chunk_len points) -- for tests */
int b200_fk20_x_ext_fft(b200_fk* fk, size_t file, uint64_t* out);
/* fk20_single.go:122-134 FK20Single: n coefficients -> n proofs, natural order */
int b200_fk20_single(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs);
/* fk20_single.go:139-172 FK20SingleDAOptimized: n2 coefficients with zero upper half -> n2 proofs */
int b200_fk20_single_da_optimized(b200_fk* fk, const uint64_t* poly, size_t n2, uint64_t* proofs);
/* fk20_single.go:176-196 DAUsingFK20: n coefficients -> 2n proofs in reverse bit order */
int b200_da_using_fk20(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs);
/* DAUsingFK20 for `batch` polynomials of n coefficients (2n proofs each, reverse bit order), host buffers / device buffers. */
int b200_da_using_fk20_batch(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint64_t* proofs);
int b200_da_using_fk20_batch_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_proofs, void* cuda_stream);
/* fk20_multi.go:58-109 FK20MultiDAOptimized: n2 coefficients (upper half zero) -> 2k proofs */
int b200_fk20_multi_da_optimized(b200_fk* fk, const uint64_t* poly, size_t n2, uint64_t* proofs);
/* fk20_multi.go:113-133 DAUsingFK20Multi: n coefficients -> 2k proofs in reverse bit order */
int b200_da_using_fk20_multi(b200_fk* fk, const uint64_t* poly, size_t n, uint64_t* proofs);

/* Headline unit of work, batched: for each of `batch` polynomials of n coefficients,
 * commitment = CommitToPoly(poly) and proofs = FK20Single(poly).  Host buffers; the copies
 * are inside the call. */
int b200_commit_fk20_batch(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint64_t* commitments,
                           uint64_t* proofs);
/* Same with DEVICE buffers (inputs resident in HBM), asynchronous on `cuda_stream`
 * (a cudaStream_t, may be NULL for the default stream). */
int b200_commit_fk20_batch_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_commitments,
                               void* d_proofs, void* cuda_stream);
/* ------------------------------------------------------------------ compressed outputs, eth blobs --
 * The callers of the path consume 48-byte compressed points (bls/bls_kilic.go:114-116 ToCompressedG1;
 * eth/helpers.go:98-103, :199-202).  These entry points normalise and compress on the device (one
 * inversion per 16 points), so a third of the bytes travels back to the host.  Device output pointers of the
 * _dev forms must be 16-byte aligned (48-byte strings are written as three 128-bit stores). */
int b200_commit_fk20_batch_compressed(b200_fk* fk, const uint64_t* polys, size_t n, size_t batch, uint8_t* commitments48,
                                      uint8_t* proofs48);
int b200_commit_fk20_batch_compressed_dev(b200_fk* fk, const void* d_polys, size_t n, size_t batch, void* d_commitments48,
                                          void* d_proofs48, void* cuda_stream);
/* ToCompressedG1 over n points: device-resident ABI points -> device 48-byte strings; host-buffer form. */
int b200_g1_compress_dev(const void* d_points, size_t n, void* d_out48, void* cuda_stream);
int b200_g1_to_compressed_batch(const uint64_t* points, size_t n, uint8_t* out48);
/* eth.BlobToKZGCommitment for `batch` blobs of n x 32 little-endian bytes (eth/helpers.go:264-273
 * BlobToPolynomial + :98-103 PolynomialToKZGCommitment).  `ks` must hold the bit-reversed Lagrange
 * setup (eth/globals.go:48).  ok[b] = 0 (and a zeroed output) where a field element is >= r
 * (bls/bignum_all.go:12-35 ValidFr), matching BlobToPolynomial's `false`. */
int b200_blob_to_kzg_commitment_batch(b200_ks* ks, const uint8_t* blobs, size_t n, size_t batch, uint8_t* out48, uint8_t* ok);
/* eth.ComputeKZGProof for a batch (eth/helpers.go:179-203, with bls.EvaluatePolyInEvaluationForm
 * bls/globals.go:106-153): polys[b] = n evaluations on the bit-reversed domain (eth/globals.go:60-67),
 * z[b] the challenge.  proofs48[b] = compressed LinCombG1(lagrange setup, (f - y) / (D - z)), y[b] = f_b(z_b)
 * (canonical, may be NULL).  ok[b] = 0 where the reference returns an error ("invalid z challenge": z in
 * the domain) or an input is not a field element.  n: power of two >= 16, <= both settings' widths. */
int b200_compute_kzg_proof_batch(b200_ks* ks, const uint64_t* polys, const uint64_t* z, size_t n, size_t batch,
                                 uint8_t* proofs48, uint64_t* y, uint8_t* ok);

/* ------------------------------------------------------------------ multi-GPU building blocks --
 * FK20 multi sharded by chunk offset: rank g computes the partial
 * hExtFFT of its offsets [off_begin, off_end) (fk20_multi.go:80-91 restricted to those files),
 * the partials (2k points each, ABI encoding, device memory) are exchanged as raw limbs over
 * NCCL -- elliptic-curve addition is not an NCCL reduction, so the "allreduce" is an all-gather
 * followed by b200_g1_sum_dev -- and b200_fk20_multi_finish_dev runs the two G1 transforms
 * (fk20_multi.go:93-104, and the reverse-bit-order of :131 when reverse_bits != 0).
 * All pointers are device pointers; calls are asynchronous on `cuda_stream`. */
int b200_fk20_multi_partial_dev(b200_fk* fk, const void* d_poly, size_t n, size_t off_begin, size_t off_end,
                                void* d_partial, void* cuda_stream);
int b200_g1_sum_dev(const void* d_parts, size_t n_parts, size_t count, void* d_out, void* cuda_stream);
int b200_fk20_multi_finish_dev(b200_fk* fk, const void* d_h_ext_fft, int reverse_bits, void* d_proofs, void* cuda_stream);
/* The same two G1 transforms spread over world = 2^s ranks that all hold the summed hExtFFT: _local leaves rank r's
 * block of k2 / world points (internal 144-byte form) after the inverse transform, the zero padding and the block-local
 * forward stages; the blocks are all-gathered in rank order; _merge runs the last s forward stages and converts. */
int b200_fk20_multi_finish_local_dev(b200_fk* fk, const void* d_h_ext_fft, size_t rank, size_t world, void* d_block, void* cuda_stream);
int b200_fk20_multi_finish_merge_dev(b200_fk* fk, const void* d_blocks, size_t world, int reverse_bits, void* d_proofs, void* cuda_stream);
/* The merge sharded as well (world^2 <= k2): after the all-gather of the blocks, rank r runs the last s forward stages only for
 * the positions [r sub, (r+1) sub), sub = k2 / world^2, of every block (d_blocks is overwritten) and leaves them in d_part
 * (k2 / world internal points); the parts are all-gathered in rank order and _assemble converts them to the proof array. */
int b200_fk20_multi_finish_merge_part_dev(b200_fk* fk, void* d_blocks, size_t rank, size_t world, void* d_part, void* cuda_stream);
int b200_fk20_multi_finish_assemble_dev(b200_fk* fk, const void* d_parts, size_t world, int reverse_bits, void* d_proofs, void* cuda_stream);
/* Partial LinCombG1 over points/scalars [begin, end) of the settings' SecretG1 (MSM sharded by
 * point range); partial sums are exchanged and added with b200_g1_sum_dev. */
int b200_commit_partial_dev(b200_ks* ks, const void* d_coeffs, size_t begin, size_t end, void* d_out, void* cuda_stream);
/* setup.go:9-26 GenerateTestingSetup, G1 half: out[i] = secret^i * G (host buffer, device compute). */
int b200_generate_testing_setup_g1(const uint64_t secret[4], size_t n, uint64_t* out);

/* Number of kernels launched by the last FK20 / commit+FK20 call made on the CALLING THREAD (bench.py
 * gpu_launches).  Kept per thread so that handles stay immutable; the fk argument is ignored. */
uint64_t b200_last_launch_count(void);
uint64_t b200_fk20_last_launch_count(const b200_fk* fk);
/* Window width (bits) of fixed-base tables built from now on: 8 (default; 384 KiB per base), 10, 12 (1.3 / 4.1 MiB
 * per base, fewer additions per look-up) or 4.  Also settable with the environment variable B200_FB_WINDOW. */
int b200_set_fixed_base_window(int bits);

/* Per-kernel-class device timing for bench.py's roofline: between begin and end every launch is
 * bracketed by CUDA events on its stream.  Classes: 0 Fr NTT, 1 G1 FFT butterfly stage,
 * 2 G1 scalar multiplication by programs / variable scalars, 3 G1 fold/add, 4 conversions and copies,
 * 5 fixed-base table look-up sums (ToeplitzPart2, commitment terms), 6 bucket MSM. */
#define B200_PROFILE_CLASSES 7
int b200_profile_begin(void);
int b200_profile_end(double ms_per_class[B200_PROFILE_CLASSES], uint64_t launches_per_class[B200_PROFILE_CLASSES]);

/* Self-test hook (tests/): run the device field/curve primitives against their portable forms
 * on `n` pseudo-random operands; mismatches[0] = total, mismatches[1 + i] = count of check i. */
int b200_selftest_field(size_t n, uint64_t seed, uint64_t mismatches[17]);
/* Integer-pipe probe (bench.py's second roofline): `threads` lanes each run 2 * iters dependent
 * 381-bit Montgomery multiplications; *ms receives the device time of that launch. */
int b200_probe_fp_mul(size_t threads, int iters, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* B200_KZG_H */
