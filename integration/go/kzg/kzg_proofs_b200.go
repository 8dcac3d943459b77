//go:build bignum_b200
// +build bignum_b200

// Siblings of kzg_single_proofs.go:17-19, 57-75 and kzg_multi_proofs.go:47-88: commitments and the G1 side of the
// checks on the device, G2 arithmetic and the pairing in the library's host code (SecretG2 was handed over by
// NewKZGSettings).  ComputeProofSingle / Multi and CommitToPolyUnoptimized keep their generic Go bodies over
// bls.LinCombG1 (the device MSM).  The *Batch / *Aggregate methods are additions for data-availability sampling:
// many samples per call, per-proof answers or one pairing for the lot.
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import (
	"crypto/rand"
	"unsafe"

	"github.com/protolambda/go-kzg/bls"
)

// KZG commitment to polynomial in evaluation form, i.e. eval = FFT(coeffs).
func CommitToEvalPoly(secretG1IFFT []bls.G1Point, eval []bls.Fr) *bls.G1Point {
	return bls.LinCombG1(secretG1IFFT, eval)
}

// KZG commitment to polynomial in coefficient form
func (ks *KZGSettings) CommitToPoly(coeffs []bls.Fr) *bls.G1Point {
	res := make([]bls.G1Point, 1)
	rc := C.b200_commit_to_poly(ks.handle, frs(coeffs), C.size_t(len(coeffs)), g1s(res))
	if rc == C.B200_ERR_LEN_MISMATCH {
		panic("runtime error: slice bounds out of range") // ks.SecretG1[:len(coeffs)]
	}
	mustB200(rc)
	return &res[0]
}

// Check a proof for a KZG commitment for an evaluation f(x) = y
func (ks *KZGSettings) CheckProofSingle(commitment *bls.G1Point, proof *bls.G1Point, x *bls.Fr, y *bls.Fr) bool {
	var ok C.int
	mustB200(C.b200_check_proof_single(ks.handle, g1s([]bls.G1Point{*commitment}), g1s([]bls.G1Point{*proof}),
		frs([]bls.Fr{*x}), frs([]bls.Fr{*y}), &ok))
	return ok == 1
}

// Check a proof for a KZG commitment for an evaluation f(x w^i) = y_i
// The ys must have a power of 2 length
func (ks *KZGSettings) CheckProofMulti(commitment *bls.G1Point, proof *bls.G1Point, x *bls.Fr, ys []bls.Fr) bool {
	var ok C.int
	rc := C.b200_check_proof_multi(ks.handle, g1s([]bls.G1Point{*commitment}), g1s([]bls.G1Point{*proof}), frs([]bls.Fr{*x}),
		frs(ys), C.size_t(len(ys)), &ok)
	if rc == C.B200_ERR_TOO_LARGE || rc == C.B200_ERR_NOT_POW2 {
		panic("ys is bad, cannot compute FFT")
	}
	if rc == C.B200_ERR_TOO_SMALL || rc == C.B200_ERR_LEN_MISMATCH {
		panic("runtime error: index out of range") // ks.SecretG2[len(ys)] / ks.SecretG1[:len(ys)]
	}
	mustB200(rc)
	return ok == 1
}

// CheckProofSingleBatch: one answer per (commitment, proof, x, y); G1 sides in one device call, pairing checks on all host cores.
func (ks *KZGSettings) CheckProofSingleBatch(commitments, proofs []bls.G1Point, xs, ys []bls.Fr) []bool {
	n := len(commitments)
	if len(proofs) != n || len(xs) != n || len(ys) != n {
		panic("CheckProofSingleBatch: length mismatch")
	}
	out := make([]bool, n)
	if n == 0 {
		return out
	}
	ok := make([]byte, n)
	mustB200(C.b200_check_proof_single_batch(ks.handle, g1s(commitments), g1s(proofs), frs(xs), frs(ys), C.size_t(n),
		(*C.uint8_t)(unsafe.Pointer(&ok[0]))))
	for i := range ok {
		out[i] = ok[i] == 1
	}
	return out
}

// randomScalars: non-zero scalars below 2^248 from crypto/rand (little-endian limbs; far below r, so always canonical)
func randomScalars(n int) []bls.Fr {
	out := make([]bls.Fr, n)
	for i := range out {
		var b [32]byte
		if _, err := rand.Read(b[:31]); err != nil {
			panic(err)
		}
		b[0] |= 1
		bls.FrFrom32(&out[i], b)
	}
	return out
}

// CheckProofSingleAggregate: all proofs with ONE pairing (random linear combination; three device MSMs).
func (ks *KZGSettings) CheckProofSingleAggregate(commitments, proofs []bls.G1Point, xs, ys []bls.Fr) bool {
	n := len(commitments)
	if len(proofs) != n || len(xs) != n || len(ys) != n {
		panic("CheckProofSingleAggregate: length mismatch")
	}
	if n == 0 {
		return true
	}
	var ok C.int
	mustB200(C.b200_check_proof_single_aggregate(ks.handle, g1s(commitments), g1s(proofs), frs(xs), frs(ys), frs(randomScalars(n)),
		C.size_t(n), &ok))
	return ok == 1
}

// CheckProofMultiAggregate: `len(xs)` samples of chunkLen values each (ys is the concatenation) with ONE pairing.
func (ks *KZGSettings) CheckProofMultiAggregate(commitments, proofs []bls.G1Point, xs, ys []bls.Fr, chunkLen uint64) bool {
	n := len(commitments)
	if len(proofs) != n || len(xs) != n || uint64(len(ys)) != uint64(n)*chunkLen {
		panic("CheckProofMultiAggregate: length mismatch")
	}
	if n == 0 {
		return true
	}
	var ok C.int
	mustB200(C.b200_check_proof_multi_aggregate(ks.handle, g1s(commitments), g1s(proofs), frs(xs), frs(ys), C.size_t(chunkLen),
		frs(randomScalars(n)), C.size_t(n), &ok))
	return ok == 1
}
