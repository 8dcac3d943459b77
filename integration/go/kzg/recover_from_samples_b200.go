//go:build bignum_b200
// +build bignum_b200

// Sibling of recover_from_samples.go:42-109.  []*bls.Fr holds Go pointers, which cgo may not pass: the shim flattens
// it into values plus a presence mask.  The device path implements zeroPolyFn = ZeroPolyViaMultiplication; any other
// zeroPolyFn runs the reference's generic body (kept in recover_from_samples_generic.go, the untouched original).
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"errors"
	"reflect"
	"unsafe"

	"github.com/protolambda/go-kzg/bls"
)

func (fs *FFTSettings) RecoverPolyFromSamples(samples []*bls.Fr, zeroPolyFn ZeroPolyFn) ([]bls.Fr, error) {
	if reflect.ValueOf(zeroPolyFn).Pointer() != reflect.ValueOf(fs.ZeroPolyViaMultiplication).Pointer() {
		return fs.recoverPolyFromSamplesGeneric(samples, zeroPolyFn)
	}
	n := len(samples)
	out := make([]bls.Fr, n)
	if n == 0 {
		return out, nil
	}
	flat := make([]bls.Fr, n)
	present := make([]byte, n)
	for i, s := range samples {
		if s != nil {
			flat[i] = *s
			present[i] = 1
		}
	}
	rc := C.b200_recover_poly_from_samples(fs.handle, frs(flat), (*C.uint8_t)(unsafe.Pointer(&present[0])), C.size_t(n), frs(out))
	switch rc {
	case C.B200_ERR_RECOVERY:
		return nil, errors.New("failed to reconstruct data correctly, changed value") // recover_from_samples.go:103-107
	case C.B200_ERR_ZERO_EVAL:
		panic("bad zero eval") // :54-58
	case C.B200_ERR_TOO_SMALL:
		panic("domain too small for requested length") // zero_poly.go:120-122
	case C.B200_ERR_NOT_POW2:
		panic("length not a power of two") // zero_poly.go:123-125
	}
	mustB200(rc)
	return out, nil
}
