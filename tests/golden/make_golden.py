#!/usr/bin/env python
"""Extract the golden vectors the reference's own tests/fixtures hold for the hot path.

Run in the BUILD container (where /root/reference is mounted):
    python tests/golden/make_golden.py
Writes (committed; /root/reference does not exist on the GPU box):
    tests/golden/reference_goldens.json   literal known-answer vectors from *_test.go
    tests/golden/trusted_setup_g1.bin     eth/trusted_setup.json: setup_G1 (4096x48 B) ||
                                          setup_G1_lagrange (4096x48 B), raw compressed bytes
Only numeric literals / fixture bytes are extracted; no reference source is copied.
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def tofr_list(path, lo, hi):
    """All decimal literals inside ToFr("...") between 1-based lines lo..hi."""
    lines = open(os.path.join(REF, path)).read().split("\n")[lo - 1:hi]
    return [m for l in lines for m in re.findall(r'ToFr\("(\d+)"\)', l)]


def main():
    g = {}
    # fft_fr_test.go:32-71 TestInvFFT: IFFT of [0..15] at scale 4
    g["inv_fft_scale4"] = {"src": "fft_fr_test.go:48-65", "input": list(range(16)),
                           "expected": tofr_list("fft_fr_test.go", 48, 65)}
    # das_extension_test.go:11-40
    g["das_ext_scale4"] = {"src": "das_extension_test.go:25-34", "input": list(range(8)),
                           "expected": tofr_list("das_extension_test.go", 25, 34)}
    # zero_poly_test.go:133-198
    src = open(os.path.join(REF, "zero_poly_test.go")).read().split("\n")[135:141]
    exists = [tok == "true" for l in src for tok in re.findall(r"true|false", l)]
    assert len(exists) == 16
    g["zero_poly_scale4"] = {"src": "zero_poly_test.go:136-141,152-169,175-192", "exists": exists,
                             "expected_eval": tofr_list("zero_poly_test.go", 152, 169),
                             "expected_poly": tofr_list("zero_poly_test.go", 175, 192)}
    # bls/globals.go:27-60 roots of unity
    g["scale2_root_of_unity"] = {"src": "bls/globals.go:27-60",
                                 "values": tofr_list("bls/globals.go", 27, 60)}
    assert len(g["scale2_root_of_unity"]["values"]) == 32
    g["modulus"] = re.search(r'ModulusStr = "(\d+)"', open(os.path.join(REF, "bls/globals.go")).read()).group(1)
    # bls/bls_test.go:11-23 TestPointCompression
    t = open(os.path.join(REF, "bls/bls_test.go")).read()
    scalar = re.search(r'SetFr\(&x, "(\d+)"\)', t).group(1)
    exp = re.search(r"expected := \[\]byte\{([^}]*)\}", t).group(1)
    g["point_compression"] = {"src": "bls/bls_test.go:13,18", "scalar": scalar,
                              "expected": [int(v) for v in exp.split(",")]}
    # bls/bls_hbls.go:23-24 generator (decimal)
    h = open(os.path.join(REF, "bls/bls_hbls.go")).read()
    g["g1_generator"] = {"src": "bls/bls_hbls.go:23-24",
                         "xy": re.findall(r'SetString\("(\d{100,})", 10\)', h)[:2]}
    # deterministic test inputs whose outputs the oracle derives
    g["fk20_single_test"] = {"src": "fk20_single_test.go:13-17", "secret": "1927409816240961209460912649124",
                             "poly": [1, 2, 3, 4, 7, 7, 7, 7, 13, 13, 13, 13, 13, 13, 13, 13],
                             "fft_scale": 5, "n2": 32}
    g["fk20_multi_test"] = {"src": "fk20_multi_test.go:12-32", "secret": "1927409816240961209460912649124",
                            "chunk_len": 16, "chunk_count": 32, "fft_scale": 10,
                            "row": [1, 2, 3, "4+i", 7, "8+i*i", 9, 10, 13, 14, 1, 15, "r-1", 1000, "r-134", 33]}
    ts = json.load(open(os.path.join(REF, "eth/trusted_setup.json")))
    g["trusted_setup"] = {"src": "eth/trusted_setup.json", "n": len(ts["setup_G1"]),
                          "roots_of_unity_first4": [str(v) for v in ts["roots_of_unity"][:4]],
                          "secret_verified": "1337"}
    # G2: generator (bls/bls_hbls.go:27-30, decimal) and the head of setup_G2 (compressed, = 1337^i * GenG2) plus the
    # entries CheckProofMulti reads for n = 8, 16 -- pins G2 compression / decompression / scalar multiplication
    g["g2_generator"] = {"src": "bls/bls_hbls.go:27-30", "x0x1y0y1": re.findall(r'D\[[01]\]\.SetString\("(\d{100,})", 10\)', h)[:4]}
    assert len(g["g2_generator"]["x0x1y0y1"]) == 4
    g["trusted_setup_g2"] = {"src": "eth/trusted_setup.json setup_G2", "n": len(ts["setup_G2"]),
                             "entries": {str(i): ts["setup_G2"][i] for i in (0, 1, 2, 3, 8, 16, 4095)}}
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(g, f, indent=1)
    with open(os.path.join(HERE, "trusted_setup_g1.bin"), "wb") as f:
        for k in ("setup_G1", "setup_G1_lagrange"):
            for hx in ts[k]:
                b = bytes.fromhex(hx)
                assert len(b) == 48
                f.write(b)
    print("wrote goldens")


if __name__ == "__main__":
    main()
