import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import go_kzg_b200 as kzg
from oracle import cref, pyref
R = pyref.R_MOD
for n, scale in ((4, 2), (4, 3), (8, 3)):
    fs, fo = kzg.FFTSettings(scale), pyref.FFTSettings(scale)
    ks = list(range(1, n + 1))
    pts = cref.g1_mul_gen(ks)
    for inv in (False, True):
        got = kzg.g1_to_compressed(fs.fft_g1(pts, inv))
        want = cref.g1_compress(cref.g1_mul_gen(fo.fft(ks, inv) if scale == (n.bit_length() - 1) else pyref.FFTSettings(n.bit_length() - 1).fft(ks, inv)))
        print(n, scale, inv, [bool((got[i] == want[i]).all()) for i in range(n)])
