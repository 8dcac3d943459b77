"""World-size-2 checks of the multi-GPU host logic on CPU (gloo): the blob partition is disjoint
and complete, the synthetic inputs of different ranks differ, and the timing reduction is the
max over ranks.  The data path itself has no collective (blob-parallel replicas)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from go_kzg_b200 import shard
from go_kzg_b200.synth import blob_polys


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        per_gpu = 3
        mine = shard.blob_range(rank, world, per_gpu)
        polys = blob_polys(per_gpu, 8, first_blob=mine.start)
        # every rank learns every rank's blob indices and a digest of its inputs
        idx = torch.tensor(list(mine), dtype=torch.int64)
        gathered = [torch.zeros_like(idx) for _ in range(world)]
        dist.all_gather(gathered, idx)
        digest = torch.from_numpy(polys.view(np.int64).reshape(per_gpu, -1).sum(axis=1))
        digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        t = shard.max_over_ranks(1.0 + rank, dist)
        rate = shard.whole_job_rate(per_gpu, world, 2, t)
        dist.barrier()
        np.save(os.path.join(out_dir, "r%d.npy" % rank),
                np.array([sorted(int(v) for g in gathered for v in g) == list(range(world * per_gpu)),
                          len({int(v) for d in digests for v in d}) == world * per_gpu,
                          t == float(world), abs(rate - world * per_gpu * 2 / world) < 1e-12], dtype=np.int64))
    finally:
        dist.destroy_process_group()


def test_blob_partition_and_time_reduction_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(os.path.join(str(tmp_path), "r%d.npy" % r)).tolist() == [1, 1, 1, 1]


def test_blob_range_edges():
    assert list(shard.blob_range(0, 1, 4)) == [0, 1, 2, 3]
    assert list(shard.blob_range(3, 8, 2)) == [6, 7]
    with pytest.raises(ValueError):
        shard.blob_range(2, 2, 1)
    assert shard.max_over_ranks(3.5) == 3.5
    assert shard.blob_seed(5) == 0xB2000005


def test_offset_and_point_ranges_partition():
    """config 5 sharding: chunk offsets / MSM points are split into contiguous, disjoint, covering
    ranges for every world size"""
    from go_kzg_b200 import multi_gpu
    for world in (1, 2, 3, 4, 8, 16, 24):
        for total in (1, 16, 4096):
            got = [i for r in range(world) for i in multi_gpu.offset_range(r, world, total)]
            assert got == list(range(total))
    assert list(multi_gpu.offset_range(1, 8, 16)) == [2, 3]
    with pytest.raises(ValueError):
        multi_gpu.point_range(4, 4, 10)


def test_block_sharded_transform_rule():
    """the G1 transforms of FK20 multi are block-sharded for power-of-two worlds only (the library refuses others)"""
    from go_kzg_b200 import multi_gpu
    assert [w for w in range(1, 20) if multi_gpu.block_sharded_transforms(w, 1 << 17)] == [2, 4, 8, 16]
    assert not multi_gpu.block_sharded_transforms(8, 8) and multi_gpu.block_sharded_transforms(4, 8)


def test_transform_sharding_rules():
    """which parts of the FK20-multi finish are spread over the ranks (multi_gpu.block_sharded_transforms / merge_sharded)"""
    from go_kzg_b200 import multi_gpu
    k2 = 1 << 17
    assert not multi_gpu.block_sharded_transforms(1, k2)
    assert all(multi_gpu.block_sharded_transforms(w, k2) and multi_gpu.merge_sharded(w, k2) for w in (2, 4, 8))
    assert not multi_gpu.block_sharded_transforms(3, k2) and not multi_gpu.merge_sharded(3, k2)     # not a power of two
    assert multi_gpu.block_sharded_transforms(8, 16) and not multi_gpu.merge_sharded(8, 16)          # 16 points: 2 per block, 0 per (block, rank)
    assert multi_gpu.merge_sharded(8, 64)
