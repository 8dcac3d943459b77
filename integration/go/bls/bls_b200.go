//go:build bignum_b200
// +build bignum_b200

// Package bls, group half of the `bignum_b200` backend.  G1 lives in libb200kzg.so (Jacobian,
// canonical limbs, Z == 0 <=> infinity, so the Go zero value is the point at infinity as
// bls/bls_kilic.go:136 and kzg_multi_proofs.go:20 rely on).  G2 and the pairing are
// verification-only and stay on kilic inside this tag (SURVEY.md section 8b).
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

	kbls "github.com/kilic/bls12-381"
)

var ZERO_G1 G1Point

var GenG1 G1Point
var GenG2 G2Point

var ZeroG1 G1Point
var ZeroG2 G2Point

func initG1G2() {
	C.b200_g1_generator(g1p(&GenG1))
	GenG2 = G2Point(*kbls.NewG2().One())
	ZeroG1 = G1Point{}
	ZeroG2 = G2Point(*kbls.NewG2().Zero())
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

func StrG1(v *G1Point) string {
	k := toKilicG1(v)
	data := kbls.NewG1().ToUncompressed(k)
	var a, b big.Int
	a.SetBytes(data[:48])
	b.SetBytes(data[48:])
	return a.String() + "\n" + b.String()
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

// ---- G2 and the pairing: kilic, unchanged from bls/bls_kilic.go:67-157 ------------------------

type G2Point kbls.PointG2

func ClearG2(x *G2Point) { (*kbls.PointG2)(x).Zero() }

func CopyG2(dst *G2Point, v *G2Point) { *dst = *v }

func kilicFr(b *Fr) *kbls.Fr {
	be := FrTo32(b)
	for i := 0; i < 16; i++ {
		be[i], be[31-i] = be[31-i], be[i]
	}
	return new(kbls.Fr).FromBytes(be[:])
}

func MulG2(dst *G2Point, a *G2Point, b *Fr) {
	kbls.NewG2().MulScalar((*kbls.PointG2)(dst), (*kbls.PointG2)(a), kilicFr(b))
}

func AddG2(dst *G2Point, a *G2Point, b *G2Point) {
	kbls.NewG2().Add((*kbls.PointG2)(dst), (*kbls.PointG2)(a), (*kbls.PointG2)(b))
}

func SubG2(dst *G2Point, a *G2Point, b *G2Point) {
	kbls.NewG2().Sub((*kbls.PointG2)(dst), (*kbls.PointG2)(a), (*kbls.PointG2)(b))
}

func NegG2(dst *G2Point) { kbls.NewG2().Neg((*kbls.PointG2)(dst), (*kbls.PointG2)(dst)) }

func StrG2(v *G2Point) string {
	data := kbls.NewG2().ToUncompressed((*kbls.PointG2)(v))
	var a, b big.Int
	a.SetBytes(data[:96])
	b.SetBytes(data[96:])
	return a.String() + "\n" + b.String()
}

func EqualG2(a *G2Point, b *G2Point) bool {
	return kbls.NewG2().Equal((*kbls.PointG2)(a), (*kbls.PointG2)(b))
}

func ToCompressedG2(p *G2Point) []byte { return kbls.NewG2().ToCompressed((*kbls.PointG2)(p)) }

func FromCompressedG2(v []byte) (*G2Point, error) {
	p, err := kbls.NewG2().FromCompressed(v)
	return (*G2Point)(p), err
}

// toKilicG1 crosses into kilic's representation through the compressed encoding (verification and
// printing only, off the hot path).
func toKilicG1(p *G1Point) *kbls.PointG1 {
	k, err := kbls.NewG1().FromCompressed(ToCompressedG1(p))
	if err != nil {
		panic(err)
	}
	return k
}

// e(a1^(-1), a2) * e(b1,  b2) = 1_T
func PairingsVerify(a1 *G1Point, a2 *G2Point, b1 *G1Point, b2 *G2Point) bool {
	pairingEngine := kbls.NewEngine()
	pairingEngine.AddPairInv(toKilicG1(a1), (*kbls.PointG2)(a2))
	pairingEngine.AddPair(toKilicG1(b1), (*kbls.PointG2)(b2))
	return pairingEngine.Check()
}

func DebugG1s(msg string, values []G1Point) {
	var out strings.Builder
	for i := range values {
		out.WriteString(fmt.Sprintf("%s %d: %s\n", msg, i, StrG1(&values[i])))
	}
	fmt.Println(out.String())
}
