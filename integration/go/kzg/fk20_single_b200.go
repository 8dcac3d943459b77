//go:build bignum_b200
// +build bignum_b200

// Sibling of fk20_single.go:59-196.  FK20Single / FK20SingleDAOptimized / DAUsingFK20 cross cgo once per polynomial
// and run the fused device pipeline (DESIGN.md section 3); ToeplitzPart2 / ToeplitzPart3 stay available as
// stand-alone methods with the reference's signatures.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"

	"github.com/protolambda/go-kzg/bls"
)

// Performs the second part of the Toeplitz matrix multiplication algorithm
func (ks *KZGSettings) ToeplitzPart2(toeplitzCoeffs []bls.Fr, xExtFFT []bls.G1Point) (hExtFFT []bls.G1Point) {
	if uint64(len(toeplitzCoeffs)) != uint64(len(xExtFFT)) {
		panic("expected toeplitz coeffs to match xExtFFT length")
	}
	hExtFFT = make([]bls.G1Point, len(xExtFFT))
	rc := C.b200_toeplitz_part2(ks.FFTSettings.handle, frs(toeplitzCoeffs), C.size_t(len(toeplitzCoeffs)), g1s(xExtFFT), C.size_t(len(xExtFFT)), g1s(hExtFFT))
	if rc == C.B200_ERR_TOO_LARGE || rc == C.B200_ERR_LEN_MISMATCH {
		panic(fmt.Errorf("FFT failed in toeplitz part 2: %s", b200Message(rc)))
	}
	mustB200(rc)
	return hExtFFT
}

// Transform back and return the first half of the vector
func (ks *KZGSettings) ToeplitzPart3(hExtFFT []bls.G1Point) []bls.G1Point {
	out := make([]bls.G1Point, len(hExtFFT)/2, len(hExtFFT)) // the reference's slice keeps capacity 2n (fk20_single.go:163)
	rc := C.b200_toeplitz_part3(ks.FFTSettings.handle, g1s(hExtFFT), C.size_t(len(hExtFFT)), g1s(out))
	if rc != C.B200_OK && rc != C.B200_ERR_CUDA && rc != C.B200_ERR_NO_DEVICE {
		panic(fmt.Errorf("toeplitz part 3 err: %s", b200Message(rc)))
	}
	mustB200(rc)
	return out
}

// Compute all n (single) proofs according to FK20 method
func (fk *FK20SingleSettings) FK20Single(polynomial []bls.Fr) []bls.G1Point {
	out := make([]bls.G1Point, len(polynomial))
	rc := C.b200_fk20_single(fk.handle, frs(polynomial), C.size_t(len(polynomial)), g1s(out))
	if rc == C.B200_ERR_LEN_MISMATCH {
		panic("expected toeplitz coeffs to match xExtFFT length") // fk20_single.go:60-62
	}
	mustB200(rc)
	return out
}

// Special version of the FK20 for the situation of data availability checks:
// The upper half of the polynomial coefficients is always 0
func (fk *FK20SingleSettings) FK20SingleDAOptimized(polynomial []bls.Fr) []bls.G1Point {
	n2 := uint64(len(polynomial))
	out := make([]bls.G1Point, n2)
	rc := C.b200_fk20_single_da_optimized(fk.handle, frs(polynomial), C.size_t(n2), g1s(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		panic(fmt.Errorf("expected input of length %d (incl half of zeroes) to not exceed precomputed settings length %d", n2, fk.MaxWidth))
	case C.B200_ERR_NOT_POW2:
		panic(fmt.Errorf("expected input length to be power of two, got %d", n2))
	case C.B200_ERR_BAD_INPUT:
		panic("bad input, second half should be zeroed")
	case C.B200_ERR_LEN_MISMATCH:
		panic("expected toeplitz coeffs to match xExtFFT length")
	}
	mustB200(rc)
	return out
}

// Computes all the KZG proofs for data availability checks. This involves sampling on the double domain
// and reordering according to reverse bit order
func (fk *FK20SingleSettings) DAUsingFK20(polynomial []bls.Fr) []bls.G1Point {
	n := uint64(len(polynomial))
	out := make([]bls.G1Point, 2*n)
	rc := C.b200_da_using_fk20(fk.handle, frs(polynomial), C.size_t(n), g1s(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		panic("expected poly contents not bigger than half the size of the FK20-single settings")
	case C.B200_ERR_NOT_POW2:
		panic("expected poly length to be power of two")
	case C.B200_ERR_LEN_MISMATCH:
		panic("expected toeplitz coeffs to match xExtFFT length")
	}
	mustB200(rc)
	return out
}

// CommitAndFK20SingleBatch is the throughput form (one cgo crossing per batch; lanes of a warp = polynomials):
// commitments[b] = CommitToPoly(polys[b]), proofs[b] = FK20Single(polys[b]); polys is batch x n, flattened.
func (fk *FK20SingleSettings) CommitAndFK20SingleBatch(polys []bls.Fr, n uint64) (commitments []bls.G1Point, proofs []bls.G1Point) {
	batch := uint64(len(polys)) / n
	commitments = make([]bls.G1Point, batch)
	proofs = make([]bls.G1Point, batch*n)
	mustB200(C.b200_commit_fk20_batch(fk.handle, frs(polys), C.size_t(n), C.size_t(batch), g1s(commitments), g1s(proofs)))
	return
}
