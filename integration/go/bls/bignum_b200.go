//go:build bignum_b200
// +build bignum_b200

// Package bls, scalar-field half of the `bignum_b200` backend: Fr arithmetic behind libb200kzg.so.
//
// Drop-in sibling of bls/bignum_kilic.go (same exported symbols, same semantics: dst may alias the
// operands, the zero value is 0).  Level-1 operations run on the host inside the library; the hot
// path crosses cgo at batch granularity in package kzg (fft_fr_b200.go and friends).
// NOT COMPILED in the build image of this repository (no Go toolchain there); kept mechanical.
package bls

/*
#cgo CFLAGS: -I${SRCDIR}/../third_party/b200kzg/include
#cgo LDFLAGS: -L${SRCDIR}/../third_party/b200kzg/lib -lb200kzg -Wl,-rpath,${SRCDIR}/../third_party/b200kzg/lib
#include "b200_kzg.h"
*/
import "C"

import (
	"crypto/rand"
	"math/big"
	"unsafe"
)

func init() {
	initGlobals()
	ClearG1(&ZERO_G1)
	initG1G2()
}

// Fr is a canonical residue in [0, r): four little-endian 64-bit limbs (== FrTo32 bytes).
// bls/bignum_kilic.go:23 keeps Montgomery form instead; the type is opaque outside this package.
type Fr [4]uint64

func frp(p *Fr) *C.uint64_t { return (*C.uint64_t)(unsafe.Pointer(p)) }

var modulusBig, _ = new(big.Int).SetString(ModulusStr, 10)

func frFromBig(dst *Fr, v *big.Int) {
	var t big.Int
	t.Mod(v, modulusBig)
	var buf [32]byte
	t.FillBytes(buf[:]) // big-endian
	for i := 0; i < 4; i++ {
		var w uint64
		for j := 0; j < 8; j++ {
			w = w<<8 | uint64(buf[32-8*(i+1)+j])
		}
		dst[i] = w
	}
}

func frToBig(v *Fr) *big.Int {
	var buf [32]byte
	for i := 0; i < 4; i++ {
		w := v[i]
		for j := 0; j < 8; j++ {
			buf[31-8*i-j] = byte(w >> (8 * j))
		}
	}
	return new(big.Int).SetBytes(buf[:])
}

// SetFr parses a decimal string (bls/bignum_kilic.go:25-29).
func SetFr(dst *Fr, v string) {
	var bv big.Int
	bv.SetString(v, 10)
	frFromBig(dst, &bv)
}

// FrFrom32 mutates the fr num. The value v is little-endian 32-bytes.
// Returns false, without modifying dst, if the value is out of range (bls/bignum_kilic.go:33-43).
func FrFrom32(dst *Fr, v [32]byte) (ok bool) {
	if !ValidFr(v) {
		return false
	}
	for i := 0; i < 4; i++ {
		var w uint64
		for j := 7; j >= 0; j-- {
			w = w<<8 | uint64(v[8*i+j])
		}
		dst[i] = w
	}
	return true
}

// FrTo32 serializes a fr number to 32 bytes. Encoded little-endian (bls/bignum_kilic.go:46-55).
func FrTo32(src *Fr) (v [32]byte) {
	for i := 0; i < 4; i++ {
		for j := 0; j < 8; j++ {
			v[8*i+j] = byte(src[i] >> (8 * j))
		}
	}
	return
}

func CopyFr(dst *Fr, v *Fr) { *dst = *v }

func AsFr(dst *Fr, i uint64) { *dst = Fr{i, 0, 0, 0} }

func FrStr(b *Fr) string {
	if b == nil {
		return "<nil>"
	}
	return frToBig(b).String()
}

func EqualOne(v *Fr) bool { return *v == Fr{1, 0, 0, 0} }

func EqualZero(v *Fr) bool { return *v == Fr{} }

func EqualFr(a *Fr, b *Fr) bool { return *a == *b }

func RandomFr() *Fr {
	v, err := rand.Int(rand.Reader, modulusBig)
	if err != nil {
		panic(err)
	}
	var out Fr
	frFromBig(&out, v)
	return &out
}

func SubModFr(dst *Fr, a, b *Fr) { C.b200_fr_sub(frp(dst), frp(a), frp(b)) }

func AddModFr(dst *Fr, a, b *Fr) { C.b200_fr_add(frp(dst), frp(a), frp(b)) }

func DivModFr(dst *Fr, a, b *Fr) { C.b200_fr_div(frp(dst), frp(a), frp(b)) }

func MulModFr(dst *Fr, a, b *Fr) { C.b200_fr_mul(frp(dst), frp(a), frp(b)) }

func InvModFr(dst *Fr, v *Fr) { C.b200_fr_inv(frp(dst), frp(v)) }

func BatchInvModFr(f []Fr) {
	if len(f) == 0 {
		return
	}
	C.b200_fr_batch_inv(frp(&f[0]), C.size_t(len(f)))
}

func EvalPolyAt(dst *Fr, p []Fr, x *Fr) {
	EvalPolyAtUnoptimized(dst, p, x)
}

// ExpModFr: square and multiply over the library's MulModFr (bls/bignum_kilic.go:129-131).
func ExpModFr(dst *Fr, v *Fr, e *big.Int) {
	var acc Fr
	AsFr(&acc, 1)
	base := *v
	for i := e.BitLen() - 1; i >= 0; i-- {
		MulModFr(&acc, &acc, &acc)
		if e.Bit(i) == 1 {
			MulModFr(&acc, &acc, &base)
		}
	}
	*dst = acc
}
