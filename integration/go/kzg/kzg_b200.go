//go:build bignum_b200
// +build bignum_b200

// Sibling of kzg.go:11-116: the settings types keep their Go fields (SecretG1 / SecretG2 are read by callers and by
// the verification code) and gain device handles holding SecretG1, xExtFFT(Files) and the fixed-base window tables.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"unsafe"

	"github.com/protolambda/go-kzg/bls"
)

type KZGSettings struct {
	*FFTSettings

	// setup values
	// [b.multiply(b.G1, pow(s, i, MODULUS)) for i in range(WIDTH+1)],
	SecretG1 []bls.G1Point
	// [b.multiply(b.G2, pow(s, i, MODULUS)) for i in range(WIDTH+1)],
	SecretG2 []bls.G2Point

	handle *C.b200_ks
}

func NewKZGSettings(fs *FFTSettings, secretG1 []bls.G1Point, secretG2 []bls.G2Point) *KZGSettings {
	if len(secretG1) != len(secretG2) {
		panic("secret list lengths don't match")
	}
	if uint64(len(secretG1)) < fs.MaxWidth {
		panic(fmt.Errorf("expected more values for secrets, MaxWidth: %d, got: %d", fs.MaxWidth, len(secretG1)))
	}
	ks := &KZGSettings{FFTSettings: fs, SecretG1: secretG1, SecretG2: secretG2}
	mustB200(C.b200_kzg_settings_new(fs.handle, g1s(secretG1), C.size_t(len(secretG1)), C.size_t(len(secretG2)), &ks.handle))
	if len(secretG2) > 0 { // the verification entry points read SecretG2[1] and SecretG2[len(ys)] from the handle
		mustB200(C.b200_kzg_settings_set_secret_g2(ks.handle, (*C.uint64_t)(unsafe.Pointer(&secretG2[0])), C.size_t(len(secretG2))))
	}
	runtime.SetFinalizer(ks, func(s *KZGSettings) { C.b200_kzg_settings_free(s.handle) })
	return ks
}

type FK20SingleSettings struct {
	*KZGSettings
	xExtFFT []bls.G1Point // filled on demand (XExtFFT); the working copy lives in HBM behind handle
	handle  *C.b200_fk
}

func fk20Panic(rc C.int) {
	switch rc {
	case C.B200_ERR_TOO_LARGE:
		panic("extended size is larger than kzg settings supports") // kzg.go:44-46, 74-76 (chunk length too large: :83-85)
	case C.B200_ERR_NOT_POW2:
		panic("extended size is not a power of two") // kzg.go:47-49, 77-79, 86-88
	case C.B200_ERR_TOO_SMALL:
		panic("extended size is too small") // kzg.go:50-52, 80-82, 89-91
	}
	mustB200(rc)
}

func NewFK20SingleSettings(ks *KZGSettings, n2 uint64) *FK20SingleSettings {
	fk := &FK20SingleSettings{KZGSettings: ks}
	fk20Panic(C.b200_fk20_single_settings_new(ks.handle, C.size_t(n2), &fk.handle))
	runtime.SetFinalizer(fk, func(s *FK20SingleSettings) { C.b200_fk20_settings_free(s.handle) })
	return fk
}

// XExtFFT returns a copy of the precomputed FFT_G1 of the extended setup vector (kzg.go:62).
func (fk *FK20SingleSettings) XExtFFT(n2 uint64) []bls.G1Point {
	if fk.xExtFFT == nil {
		fk.xExtFFT = make([]bls.G1Point, n2)
		mustB200(C.b200_fk20_x_ext_fft(fk.handle, 0, g1s(fk.xExtFFT)))
	}
	return fk.xExtFFT
}

type FK20MultiSettings struct {
	*KZGSettings
	chunkLen uint64
	n2       uint64
	handle   *C.b200_fk
}

func NewFK20MultiSettings(ks *KZGSettings, n2 uint64, chunkLen uint64) *FK20MultiSettings {
	if chunkLen > n2/2 && n2 <= ks.MaxWidth && bls.IsPowerOfTwo(n2) && n2 >= 2 {
		panic("chunk length is too large") // kzg.go:83-85
	}
	fk := &FK20MultiSettings{KZGSettings: ks, chunkLen: chunkLen, n2: n2}
	fk20Panic(C.b200_fk20_multi_settings_new(ks.handle, C.size_t(n2), C.size_t(chunkLen), &fk.handle))
	runtime.SetFinalizer(fk, func(s *FK20MultiSettings) { C.b200_fk20_settings_free(s.handle) })
	return fk
}

// NewFK20MultiSettingsFromCache adopts xExtFFT files exported earlier (XExtFFTFiles), skipping the chunkLen G1
// transforms of kzg.go:101-114; the caller keys its cache by the setup.
func NewFK20MultiSettingsFromCache(ks *KZGSettings, n2 uint64, chunkLen uint64, files []bls.G1Point) *FK20MultiSettings {
	fk := &FK20MultiSettings{KZGSettings: ks, chunkLen: chunkLen, n2: n2}
	fk20Panic(C.b200_fk20_settings_new_from_x_ext_fft(ks.handle, C.size_t(n2), C.size_t(chunkLen), 0, C.size_t(chunkLen), g1s(files), &fk.handle))
	runtime.SetFinalizer(fk, func(s *FK20MultiSettings) { C.b200_fk20_settings_free(s.handle) })
	return fk
}

// XExtFFTFiles exports all files (chunkLen x n2/chunkLen points, concatenated) for such a cache.
func (fk *FK20MultiSettings) XExtFFTFiles() []bls.G1Point {
	k2 := fk.n2 / fk.chunkLen
	out := make([]bls.G1Point, fk.chunkLen*k2)
	for i := uint64(0); i < fk.chunkLen; i++ {
		mustB200(C.b200_fk20_x_ext_fft(fk.handle, C.size_t(i), g1s(out[i*k2:(i+1)*k2])))
	}
	return out
}
