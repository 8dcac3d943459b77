//go:build bignum_b200
// +build bignum_b200

// Sibling of fk20_multi.go:58-133.  (FK20Multi, fk20_multi.go:25-54, indexes its files out of range for chunkLen > 1 in
// the reference and has no caller; it is not provided.)
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"

	"github.com/protolambda/go-kzg/bls"
)

// FK20 multi-proof method, optimized for dava availability where the top half of polynomial
// coefficients == 0
func (ks *FK20MultiSettings) FK20MultiDAOptimized(polynomial []bls.Fr) []bls.G1Point {
	n2 := uint64(len(polynomial))
	out := make([]bls.G1Point, n2/ks.chunkLen)
	rc := C.b200_fk20_multi_da_optimized(ks.handle, frs(polynomial), C.size_t(n2), g1s(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		panic(fmt.Errorf("KZGSettings are set to MaxWidth %d but got polynomial of length %d", ks.MaxWidth, n2))
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
func (ks *FK20MultiSettings) DAUsingFK20Multi(polynomial []bls.Fr) []bls.G1Point {
	n := uint64(len(polynomial))
	out := make([]bls.G1Point, 2*n/ks.chunkLen)
	rc := C.b200_da_using_fk20_multi(ks.handle, frs(polynomial), C.size_t(n), g1s(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		panic("expected poly contents not bigger than half the size of the FK20-multi settings")
	case C.B200_ERR_NOT_POW2:
		panic("expected poly length to be power of two")
	case C.B200_ERR_LEN_MISMATCH:
		panic("expected toeplitz coeffs to match xExtFFT length")
	}
	mustB200(rc)
	return out
}
