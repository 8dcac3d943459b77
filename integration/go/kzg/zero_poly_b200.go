//go:build bignum_b200
// +build bignum_b200

// Sibling of zero_poly.go:116-217.  The device computes the same polynomial by an exact route of its own (product
// tree over monic factors / direct evaluation); the sizes at which the reference panics are replayed by the library.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"unsafe"

	"github.com/protolambda/go-kzg/bls"
)

type ZeroPolyFn func(missingIndices []uint64, length uint64) ([]bls.Fr, []bls.Fr)

func (fs *FFTSettings) ZeroPolyViaMultiplication(missingIndices []uint64, length uint64) ([]bls.Fr, []bls.Fr) {
	zeroEval := make([]bls.Fr, length)
	zeroPoly := make([]bls.Fr, length)
	if len(missingIndices) == 0 {
		return zeroEval, zeroPoly
	}
	rc := C.b200_zero_poly_via_multiplication(fs.handle, (*C.uint64_t)(unsafe.Pointer(&missingIndices[0])),
		C.size_t(len(missingIndices)), C.size_t(length), frs(zeroEval), frs(zeroPoly))
	switch rc {
	case C.B200_ERR_TOO_SMALL:
		panic("domain too small for requested length")
	case C.B200_ERR_NOT_POW2:
		panic("length not a power of two")
	case C.B200_ERR_BAD_INPUT:
		panic("expected output smaller or equal to input length") // zero_poly.go:133 / 190 / 207-209 size panics
	}
	mustB200(rc)
	return zeroEval, zeroPoly
}
