//go:build bignum_b200
// +build bignum_b200

// Sibling of fft.go:34-61 and fft_fr.go:55-105: FFTSettings keeps its exported fields (callers read
// ExpandedRootsOfUnity directly, e.g. eth/globals.go and the tests) and gains the device handle.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"
	"runtime"

	"github.com/protolambda/go-kzg/bls"
)

type FFTSettings struct {
	MaxWidth uint64
	// the generator used to get all roots of unity
	RootOfUnity *bls.Fr
	// domain, starting and ending with 1 (duplicate!)
	ExpandedRootsOfUnity []bls.Fr
	// reverse domain, same as inverse values of domain. Also starting and ending with 1.
	ReverseRootsOfUnity []bls.Fr

	handle *C.b200_fs // device-side tables (twiddles, twiddle programs, coset shifts)
}

func NewFFTSettings(maxScale uint8) *FFTSettings {
	width := uint64(1) << maxScale
	fs := &FFTSettings{
		MaxWidth:             width,
		RootOfUnity:          &bls.Scale2RootOfUnity[maxScale],
		ExpandedRootsOfUnity: make([]bls.Fr, width+1),
		ReverseRootsOfUnity:  make([]bls.Fr, width+1),
	}
	mustB200(C.b200_fft_settings_new(C.uint8_t(maxScale), &fs.handle))
	mustB200(C.b200_fs_roots(fs.handle, 0, frs(fs.ExpandedRootsOfUnity)))
	mustB200(C.b200_fs_roots(fs.handle, 1, frs(fs.ReverseRootsOfUnity)))
	runtime.SetFinalizer(fs, func(s *FFTSettings) { C.b200_fft_settings_free(s.handle) })
	return fs
}

func (fs *FFTSettings) FFT(vals []bls.Fr, inv bool) ([]bls.Fr, error) {
	n := uint64(len(vals))
	if n > fs.MaxWidth {
		return nil, fmt.Errorf("got %d values but only have %d roots of unity", n, fs.MaxWidth)
	}
	out := make([]bls.Fr, nextPowOf2(n))
	rc := C.b200_fft_fr(fs.handle, frs(vals), C.size_t(n), cbool(inv), frs(out))
	if rc == C.B200_ERR_TOO_LARGE {
		return nil, fmt.Errorf("got %d values but only have %d roots of unity", n, fs.MaxWidth)
	}
	mustB200(rc)
	return out, nil
}

func (fs *FFTSettings) InplaceFFT(vals []bls.Fr, out []bls.Fr, inv bool) error {
	n := uint64(len(vals))
	rc := C.b200_inplace_fft_fr(fs.handle, frs(vals), C.size_t(n), cbool(inv), frs(out))
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		return fmt.Errorf("got %d values but only have %d roots of unity", n, fs.MaxWidth)
	case C.B200_ERR_NOT_POW2:
		return fmt.Errorf("got %d values but not a power of two", n)
	case C.B200_ERR_BAD_INPUT:
		panic("runtime error: integer divide by zero") // fft_fr.go:89,100 with n == 0
	}
	mustB200(rc)
	return nil
}

func cbool(b bool) C.int {
	if b {
		return 1
	}
	return 0
}
