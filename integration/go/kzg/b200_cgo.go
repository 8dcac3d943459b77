//go:build bignum_b200
// +build bignum_b200

// Package kzg under the `bignum_b200` tag: cgo plumbing shared by the tag-gated siblings of the batch methods
// (fft_b200.go, fft_g1_b200.go, kzg_b200.go, fk20_single_b200.go, fk20_multi_b200.go, das_extension_b200.go,
// zero_poly_b200.go, recover_from_samples_b200.go, kzg_proofs_b200.go).  The original files keep their bodies and gain
// `&& !bignum_b200` in their build constraint (integration/go/README.md lists them).  Go signatures and result
// semantics are those of the reference.  NOT COMPILED in this repository's build image (no Go toolchain there).
package kzg

/*
#cgo CFLAGS: -I${SRCDIR}/third_party/b200kzg/include
#cgo LDFLAGS: -L${SRCDIR}/third_party/b200kzg/lib -lb200kzg -Wl,-rpath,${SRCDIR}/third_party/b200kzg/lib
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"
	"unsafe"

	"github.com/protolambda/go-kzg/bls"
)

func frs(v []bls.Fr) *C.uint64_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.uint64_t)(unsafe.Pointer(&v[0]))
}

func g1s(v []bls.G1Point) *C.uint64_t {
	if len(v) == 0 {
		return nil
	}
	return (*C.uint64_t)(unsafe.Pointer(&v[0]))
}

func b200Message(rc C.int) string {
	return fmt.Sprintf("b200kzg: %s (%s)", C.GoString(C.b200_strerror(rc)), C.GoString(C.b200_last_cuda_error()))
}

// mustB200 panics on the statuses that have no counterpart in the reference (CUDA failures, no device: there is no
// CPU fallback) and on anything a caller did not map explicitly.
func mustB200(rc C.int) {
	if rc != C.B200_OK {
		panic(b200Message(rc))
	}
}
