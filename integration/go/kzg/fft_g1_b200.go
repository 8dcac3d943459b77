//go:build bignum_b200
// +build bignum_b200

// Sibling of fft_g1.go:58-94.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"

	"github.com/protolambda/go-kzg/bls"
)

func (fs *FFTSettings) FFTG1(vals []bls.G1Point, inv bool) ([]bls.G1Point, error) {
	n := uint64(len(vals))
	out := make([]bls.G1Point, n)
	rc := C.b200_fft_g1(fs.handle, g1s(vals), C.size_t(n), cbool(inv), g1s(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		return nil, fmt.Errorf("got %d values but only have %d roots of unity", n, fs.MaxWidth)
	case C.B200_ERR_NOT_POW2:
		return nil, fmt.Errorf("got %d values but not a power of two", n)
	case C.B200_ERR_BAD_INPUT:
		panic("runtime error: integer divide by zero") // fft_g1.go:78,88 with n == 0
	}
	mustB200(rc)
	return out, nil
}

// DASFFTExtensionG1 is the G1 form of DASFFTExtension that fk20_multi.go:96 leaves as a TODO (in place).
func (fs *FFTSettings) DASFFTExtensionG1(vals []bls.G1Point) {
	rc := C.b200_das_fft_extension_g1(fs.handle, g1s(vals), C.size_t(len(vals)))
	if rc == C.B200_ERR_TOO_SMALL {
		panic("domain too small for extending requested values")
	}
	if rc == C.B200_ERR_BAD_INPUT {
		panic("bad usage")
	}
	mustB200(rc)
}
