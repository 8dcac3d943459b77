#!/usr/bin/env python
"""Multi-GPU check of the sharded FK20-multi / commitment paths (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded.py
Every rank computes the sharded result and compares it with its own single-GPU result."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import go_kzg_b200 as kzg                      # noqa: E402
from go_kzg_b200 import multi_gpu              # noqa: E402
from go_kzg_b200.synth import random_fr_limbs  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kzg.lib().b200_set_device(local)
    rank, world = dist.get_rank(), dist.get_world_size()
    scale = int(os.environ.get("SHARD_SCALE", "12"))          # n2 = 2^scale
    l = 16
    n = 1 << (scale - 1)
    secret = 1927409816240961209460912649124
    setup = kzg.generate_testing_setup_g1(secret, 2 * n)
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, setup)
    fk = kzg.FK20MultiSettings(ks, 2 * n, l)
    poly = random_fr_limbs(n, 99)
    want = fk.da_using_fk20_multi(poly)
    dist.barrier()
    t0 = time.perf_counter()
    got = multi_gpu.da_using_fk20_multi_sharded(fk, poly, dist)
    dt = time.perf_counter() - t0
    ok1 = np.array_equal(kzg.g1_to_compressed(got[:64]), kzg.g1_to_compressed(want[:64])) and \
        all(kzg.lib().b200_g1_equal(got[i].ctypes.data, want[i].ctypes.data) == 1 for i in range(0, got.shape[0], max(1, got.shape[0] // 97)))
    c_want = ks.commit_to_poly(poly)
    c_got = multi_gpu.commit_to_poly_sharded(ks, poly, dist)
    ok2 = kzg.lib().b200_g1_equal(c_got.ctypes.data, c_want.ctypes.data) == 1
    flags = torch.tensor([int(ok1), int(ok2)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded FK20 multi n=%d chunk=%d over %d GPUs: proofs %s, commitment %s, %.3f s" %
              (n, l, world, "OK" if flags[0].item() else "MISMATCH", "OK" if flags[1].item() else "MISMATCH", dt), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flags.min().item() == 1 else 1)


if __name__ == "__main__":
    main()
