//go:build bignum_b200
// +build bignum_b200

// Siblings of kzg_single_proofs.go:17-19, 57-75 and kzg_multi_proofs.go:47-88: commitments and the G1 side of the
// checks on the device, G2 arithmetic and the pairing on kilic (bls.PairingsVerify).  ComputeProofSingle / Multi and
// CommitToPolyUnoptimized keep their generic Go bodies over bls.LinCombG1 (the device MSM).
package kzg

/*
#include "b200_kzg.h"
*/
import "C"

import "github.com/protolambda/go-kzg/bls"

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
	var xG2 bls.G2Point
	bls.MulG2(&xG2, &bls.GenG2, x)
	var sMinuxX bls.G2Point
	bls.SubG2(&sMinuxX, &ks.SecretG2[1], &xG2)
	// [commitment - y]_1 on the device (kzg_single_proofs.go:63-67)
	in := []bls.G1Point{*commitment}
	ys := []bls.Fr{*y}
	out := make([]bls.G1Point, 1)
	mustB200(C.b200_check_proof_single_g1_batch(g1s(in), frs(ys), 1, g1s(out)))
	return bls.PairingsVerify(&out[0], &bls.GenG2, proof, &sMinuxX)
}

// Check a proof for a KZG commitment for an evaluation f(x w^i) = y_i
// The ys must have a power of 2 length
func (ks *KZGSettings) CheckProofMulti(commitment *bls.G1Point, proof *bls.G1Point, x *bls.Fr, ys []bls.Fr) bool {
	in := []bls.G1Point{*commitment}
	xs := []bls.Fr{*x}
	out := make([]bls.G1Point, 1)
	xPow := make([]bls.Fr, 1)
	rc := C.b200_check_proof_multi_g1_batch(ks.handle, g1s(in), frs(xs), frs(ys), C.size_t(len(ys)), 1, g1s(out), frs(xPow))
	if rc == C.B200_ERR_TOO_LARGE || rc == C.B200_ERR_NOT_POW2 {
		panic("ys is bad, cannot compute FFT")
	}
	mustB200(rc)
	// [x^n]_2, [s^n - x^n]_2 (kzg_multi_proofs.go:72-76)
	var xn2 bls.G2Point
	bls.MulG2(&xn2, &bls.GenG2, &xPow[0])
	var xnMinusYn bls.G2Point
	bls.SubG2(&xnMinusYn, &ks.SecretG2[len(ys)], &xn2)
	return bls.PairingsVerify(&out[0], &bls.GenG2, proof, &xnMinusYn)
}
