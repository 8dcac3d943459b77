//go:build bignum_b200
// +build bignum_b200

// Package bls, group half of the `bignum_b200` backend.  G1 lives in libb200kzg.so (Jacobian,
// canonical limbs, Z == 0 <=> infinity, so the Go zero value is the point at infinity as
// bls/bls_kilic.go:136 and kzg_multi_proofs.go:20 rely on).  G2 and the pairing are the library's
// host code (go_kzg_b200/csrc/pairing.h): this tag does not import kilic at all.
// Drop-in sibling of bls/bls_kilic.go.  NOT COMPILED in this repository's build image.
package bls

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math/big"
	"strings"
	"unsafe"
)

var ZERO_G1 G1Point

var GenG1 G1Point
var GenG2 G2Point

var ZeroG1 G1Point
var ZeroG2 G2Point

func initG1G2() {
	C.b200_g1_generator(g1p(&GenG1))
	C.b200_g2_generator(g2p(&GenG2))
	ZeroG1 = G1Point{}
	ZeroG2 = G2Point{}
}

// G1Point: Jacobian X, Y, Z, each six little-endian 64-bit limbs of a canonical residue mod p.
type G1Point struct{ X, Y, Z [6]uint64 }

func g1p(p *G1Point) *C.uint64_t { return (*C.uint64_t)(unsafe.Pointer(p)) }

func must(rc C.int) {
	if rc != C.B200_OK {
		panic(fmt.Errorf("b200kzg: %s (%s)", C.GoString(C.b200_strerror(rc)), C.GoString(C.b200_last_cuda_error())))
	}
}

// Status exposes the library's status codes to package kzg, which maps them to the reference's error / panic split.
type Status int

func ClearG1(x *G1Point) { *x = G1Point{} }

func CopyG1(dst *G1Point, v *G1Point) { *dst = *v }

func MulG1(dst *G1Point, a *G1Point, b *Fr) { C.b200_g1_mul(g1p(dst), g1p(a), frp(b)) }

func AddG1(dst *G1Point, a *G1Point, b *G1Point) { C.b200_g1_add(g1p(dst), g1p(a), g1p(b)) }

func SubG1(dst *G1Point, a *G1Point, b *G1Point) { C.b200_g1_sub(g1p(dst), g1p(a), g1p(b)) }

func NegG1(dst *G1Point) { C.b200_g1_neg(g1p(dst)) }

func EqualG1(a *G1Point, b *G1Point) bool { return C.b200_g1_equal(g1p(a), g1p(b)) == 1 }

func ToCompressedG1(p *G1Point) []byte {
	out := make([]byte, 48)
	C.b200_g1_to_compressed((*C.uint8_t)(unsafe.Pointer(&out[0])), g1p(p))
	return out
}

func FromCompressedG1(v []byte) (*G1Point, error) {
	if len(v) != 48 {
		return nil, errors.New("input string should be equal or larger than 48")
	}
	var p G1Point
	if C.b200_g1_from_compressed(g1p(&p), (*C.uint8_t)(unsafe.Pointer(&v[0]))) != C.B200_OK {
		return nil, errors.New("invalid compressed G1 point") // flags, x >= p, off the curve or outside the subgroup
	}
	return &p, nil
}

// FromCompressedG1Batch decodes len(v)/48 points on the device (setup files: eth/globals.go:33-49).
func FromCompressedG1Batch(v []byte) ([]G1Point, error) {
	n := len(v) / 48
	out := make([]G1Point, n)
	if n == 0 {
		return out, nil
	}
	rc := C.b200_g1_from_compressed_batch((*C.uint8_t)(unsafe.Pointer(&v[0])), C.size_t(n), g1p(&out[0]), nil)
	if rc == C.B200_ERR_BAD_INPUT {
		return nil, errors.New("invalid compressed G1 point")
	}
	must(rc)
	return out, nil
}

// limbsToBig: six little-endian 64-bit limbs -> integer
func limbsToBig(l []uint64) *big.Int {
	var be [48]byte
	for i, w := range l {
		for j := 0; j < 8; j++ {
			be[47-8*i-j] = byte(w >> (8 * uint(j)))
		}
	}
	return new(big.Int).SetBytes(be[:])
}

func StrG1(v *G1Point) string {
	var xy [12]uint64
	C.b200_g1_to_affine(g1p(v), (*C.uint64_t)(unsafe.Pointer(&xy[0])))
	return limbsToBig(xy[:6]).String() + "\n" + limbsToBig(xy[6:]).String()
}

// LinCombG1: device MSM (Pippenger bucket method).  Panics on a length mismatch, empty input gives
// infinity (bls/bls_kilic.go:132-150, bls/bls_test.go:69-77).
func LinCombG1(numbers []G1Point, factors []Fr) *G1Point {
	if len(numbers) != len(factors) {
		panic("got LinCombG1 numbers/factors length mismatch")
	}
	var out G1Point
	if len(numbers) == 0 {
		return &out
	}
	must(C.b200_g1_lincomb(g1p(&numbers[0]), frp(&factors[0]), C.size_t(len(numbers)), g1p(&out)))
	return &out
}

// ---- G2 and the pairing: host code of the library (bls/bls_kilic.go:67-157) -------------------

// G2Point: Jacobian X, Y, Z over Fp2, each coordinate (c0, c1) of six little-endian 64-bit limbs; the zero value is infinity.
type G2Point struct{ X, Y, Z [2][6]uint64 }

func g2p(p *G2Point) *C.uint64_t { return (*C.uint64_t)(unsafe.Pointer(p)) }

func ClearG2(x *G2Point) { *x = G2Point{} }

func CopyG2(dst *G2Point, v *G2Point) { *dst = *v }

func MulG2(dst *G2Point, a *G2Point, b *Fr) { C.b200_g2_mul(g2p(dst), g2p(a), frp(b)) }

func AddG2(dst *G2Point, a *G2Point, b *G2Point) { C.b200_g2_add(g2p(dst), g2p(a), g2p(b)) }

func SubG2(dst *G2Point, a *G2Point, b *G2Point) { C.b200_g2_sub(g2p(dst), g2p(a), g2p(b)) }

func NegG2(dst *G2Point) { C.b200_g2_neg(g2p(dst)) }

func StrG2(v *G2Point) string {
	// kilic's ToUncompressed lays a G2 point out as x.c1 || x.c0 || y.c1 || y.c0; StrG2 prints the two 96-byte halves
	var xy [24]uint64
	C.b200_g2_to_affine(g2p(v), (*C.uint64_t)(unsafe.Pointer(&xy[0])))
	join := func(c1, c0 []uint64) *big.Int {
		hi := limbsToBig(c1)
		return hi.Lsh(hi, 384).Or(hi, limbsToBig(c0))
	}
	return join(xy[6:12], xy[0:6]).String() + "\n" + join(xy[18:24], xy[12:18]).String()
}

func EqualG2(a *G2Point, b *G2Point) bool { return C.b200_g2_equal(g2p(a), g2p(b)) == 1 }

func ToCompressedG2(p *G2Point) []byte {
	out := make([]byte, 96)
	C.b200_g2_to_compressed((*C.uint8_t)(unsafe.Pointer(&out[0])), g2p(p))
	return out
}

func FromCompressedG2(v []byte) (*G2Point, error) {
	if len(v) != 96 {
		return nil, errors.New("input string should be equal or larger than 96")
	}
	var p G2Point
	if C.b200_g2_from_compressed(g2p(&p), (*C.uint8_t)(unsafe.Pointer(&v[0]))) != C.B200_OK {
		return nil, errors.New("invalid compressed G2 point") // flags, coordinate >= p, off the curve or outside the subgroup
	}
	return &p, nil
}

// e(a1^(-1), a2) * e(b1,  b2) = 1_T
func PairingsVerify(a1 *G1Point, a2 *G2Point, b1 *G1Point, b2 *G2Point) bool {
	var ok C.int
	must(C.b200_pairings_verify(g1p(a1), g2p(a2), g1p(b1), g2p(b2), &ok))
	return ok == 1
}

func DebugG1s(msg string, values []G1Point) {
	var out strings.Builder
	for i := range values {
		out.WriteString(fmt.Sprintf("%s %d: %s\n", msg, i, StrG1(&values[i])))
	}
	fmt.Println(out.String())
}
