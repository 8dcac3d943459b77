//go:build bignum_b200
// +build bignum_b200

// Sibling of das_extension.go:71-84.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import "github.com/protolambda/go-kzg/bls"

// Takes vals as input, the values of the even indices.
// Then computes the values for the odd indices, which combined would make the right half of coefficients zero.
// Warning: the odd results are written back to the vals slice.
func (fs *FFTSettings) DASFFTExtension(vals []bls.Fr) {
	rc := C.b200_das_fft_extension(fs.handle, frs(vals), C.size_t(len(vals)))
	if rc == C.B200_ERR_TOO_SMALL {
		panic("domain too small for extending requested values")
	}
	if rc == C.B200_ERR_BAD_INPUT {
		panic("bad usage")
	}
	mustB200(rc)
}
