"""b200-kzg: Blackwell-native KZG / FFT engine behind go-kzg's bls/ build-tag seam.

The product is the C-ABI shared library (include/b200_kzg.h, built from csrc/ by build.py).
This package is the thin host-side mirror of the reference's Go API over that ABI
(FFTSettings / KZGSettings / FK20SingleSettings / FK20MultiSettings, same names, argument
meaning and error behaviour), used by the parity tests and the benchmark.  It never falls back
to a CPU implementation: if the library or a CUDA device is missing, calls raise.
"""
from .kzg import (  # noqa: F401
    B200Error,
    KZGError,
    KZGPanic,
    FFTSettings,
    KZGSettings,
    FK20SingleSettings,
    FK20MultiSettings,
    lib,
    lib_path,
    fr_from_ints,
    fr_to_ints,
    g1_to_compressed,
    g1_to_compressed_device,
    g1_from_compressed,
    g1_from_compressed_device,
    g1_marshal_text,
    g1_unmarshal_text,
    bit_reversal_permutation,
    load_trusted_setup,
    check_proof_single_g1,
    lincomb_g1,
    g1_mul_many,
    generate_testing_setup_g1,
    generate_testing_setup_g2,
    g2_generator,
    g2_add,
    g2_sub,
    g2_neg,
    g2_mul,
    g2_equal,
    g2_to_compressed,
    g2_from_compressed,
    pairings_verify,
    pairing,
    R_MOD,
)
