#!/usr/bin/env python
"""One rank's share of the two G1 transforms of config 5 (k2 = 2^17 points, world = 8), emulated on ONE GPU: the local
part (3 half stages + block-local inverse / forward transforms), the rank's merge part and the assembly, with CUDA-event
times and the per-class profile -- what bounds the 8-GPU number.   PYTHONPATH=. python tools/rank_probe.py [world]"""
import ctypes as C
import sys

import numpy as np
import torch

import go_kzg_b200 as kzg

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
k2 = 1 << 17
L = kzg.lib()
fs = kzg.FFTSettings(17)
ks = kzg.KZGSettings(fs, kzg.generate_testing_setup_g1(1337, k2))
fk = kzg.FK20SingleSettings(ks, k2)          # chunk length 1: the finish entry points only depend on k2 = n2 / chunk_len
src = torch.from_numpy(kzg.generate_testing_setup_g1(7, k2).view(np.int64)).cuda()
blk = k2 // world
d_block = torch.zeros((blk, 18), dtype=torch.int64, device="cuda")
blocks = torch.zeros((world, blk, 18), dtype=torch.int64, device="cuda")
d_mine = torch.zeros((k2 // world, 18), dtype=torch.int64, device="cuda")
parts2 = torch.zeros((world, k2 // world, 18), dtype=torch.int64, device="cuda")
d_out = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
names = ["fr_ntt", "g1_fft_stage", "g1_mul", "g1_fold", "misc", "g1_lookup", "g1_msm"]


def timed(label, fn):
    fn()
    torch.cuda.synchronize()
    ms = (C.c_double * 7)()
    cnt = (C.c_uint64 * 7)()
    L.b200_profile_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    L.b200_profile_end(ms, cnt)
    print("%-28s %7.2f ms   " % (label, e0.elapsed_time(e1)) + "  ".join("%s %.2f ms / %d" % (n, ms[i], cnt[i]) for i, n in enumerate(names) if cnt[i]))


for mode in (1, 0):
    L.b200_set_latency_mode(mode)
    print("latency mode", mode, " world", world)
    timed("finish_local (rank 3)", lambda: L.b200_fk20_multi_finish_local_dev(fk.h, src.data_ptr(), 3, world, d_block.data_ptr(), sp))
    timed("merge_part (rank 3)", lambda: L.b200_fk20_multi_finish_merge_part_dev(fk.h, blocks.data_ptr(), 3, world, d_mine.data_ptr(), sp))
    timed("assemble", lambda: L.b200_fk20_multi_finish_assemble_dev(fk.h, parts2.data_ptr(), world, 1, d_out.data_ptr(), sp))
L.b200_set_latency_mode(1)
