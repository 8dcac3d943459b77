"""The two real exchange steps of the path on several GPUs:

* FK20 multi sharded by chunk offset (config 5): every rank computes the partial hExtFFT of its
  offsets, the partials are all-gathered as raw limbs (NCCL has no elliptic-curve reduction, so the
  "G1 allreduce" is all-gather + a local add kernel); the two G1 transforms are then block-sharded as
  well (power-of-two worlds): block-local stages, one all-gather of the blocks, the last log2(world) stages;
* LinCombG1 / CommitToPoly sharded by point range: partial sums (144 B each) are all-gathered and
  added locally.

One process per GPU; `torch.distributed` (backend nccl) carries the all-gather, torch owns the
device buffers.  The compute is the C ABI of libb200kzg.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import kzg


def offset_range(rank: int, world: int, chunk_len: int) -> range:
    """Chunk offsets owned by `rank`: contiguous, disjoint, covering [0, chunk_len)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    lo = rank * chunk_len // world
    hi = (rank + 1) * chunk_len // world
    return range(lo, hi)


def point_range(rank: int, world: int, n: int) -> range:
    """Point / scalar indices of an MSM owned by `rank`."""
    return offset_range(rank, world, n)


def block_sharded_transforms(world: int, k2: int) -> bool:
    """Whether the two size-k2 G1 transforms of FK20 multi are spread over the ranks (b200_fk20_multi_finish_local_dev /
    _merge_dev): power-of-two worlds with at least two points per block; otherwise every rank runs them in full."""
    return world > 1 and world & (world - 1) == 0 and 2 * world <= k2


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    kzg._raise(rc, what=what)


def da_using_fk20_multi_sharded(fk: "kzg.FK20MultiSettings", poly: np.ndarray, dist=None) -> np.ndarray:
    """DAUsingFK20Multi (fk20_multi.go:113-133) with the chunk offsets spread over the ranks of the
    default process group.  Every rank returns the full proof array (2k points, reverse bit order)."""
    import torch
    L = kzg.lib()
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    p = np.ascontiguousarray(poly, dtype=np.uint64).reshape(-1, 4)
    n = p.shape[0]
    k2 = 2 * n // fk.chunk_len
    d_poly = torch.from_numpy(p.view(np.int64)).cuda()
    d_part = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    mine = offset_range(rank, world, fk.chunk_len)
    sp = _stream_ptr(torch)
    _check(L.b200_fk20_multi_partial_dev(fk.h, d_poly.data_ptr(), n, mine.start, mine.stop, d_part.data_ptr(), sp), "FK20 multi partial")
    if world > 1:
        parts = torch.zeros((world, k2, 18), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(parts, d_part)
        d_sum = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
        _check(L.b200_g1_sum_dev(parts.data_ptr(), world, k2, d_sum.data_ptr(), sp), "G1 sum")
    else:
        d_sum = d_part
    d_out = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    if block_sharded_transforms(world, k2):
        # the two G1 transforms, block-sharded: local stages, one all-gather of the blocks, the last log2(world) stages
        d_block = torch.zeros((k2 // world, 18), dtype=torch.int64, device="cuda")
        _check(L.b200_fk20_multi_finish_local_dev(fk.h, d_sum.data_ptr(), rank, world, d_block.data_ptr(), sp), "FK20 multi finish (local)")
        blocks = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(blocks, d_block)
        _check(L.b200_fk20_multi_finish_merge_dev(fk.h, blocks.data_ptr(), world, 1, d_out.data_ptr(), sp), "FK20 multi finish (merge)")
    else:
        _check(L.b200_fk20_multi_finish_dev(fk.h, d_sum.data_ptr(), 1, d_out.data_ptr(), sp), "FK20 multi finish")
    torch.cuda.synchronize()
    return d_out.cpu().numpy().view(np.uint64)


def commit_to_poly_sharded(ks: "kzg.KZGSettings", coeffs: np.ndarray, dist=None) -> np.ndarray:
    """CommitToPoly (kzg_single_proofs.go:17-19) with the MSM sharded by point range."""
    import torch
    L = kzg.lib()
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    n = c.shape[0]
    d_c = torch.from_numpy(c.view(np.int64)).cuda()
    d_part = torch.zeros((1, 18), dtype=torch.int64, device="cuda")
    mine = point_range(rank, world, n)
    sp = _stream_ptr(torch)
    _check(L.b200_commit_partial_dev(ks.h, d_c.data_ptr(), mine.start, mine.stop, d_part.data_ptr(), sp), "commit partial")
    if world > 1:
        parts = torch.zeros((world, 1, 18), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(parts, d_part)
        d_sum = torch.zeros((1, 18), dtype=torch.int64, device="cuda")
        _check(L.b200_g1_sum_dev(parts.data_ptr(), world, 1, d_sum.data_ptr(), sp), "G1 sum")
    else:
        d_sum = d_part
    torch.cuda.synchronize()
    return d_sum.cpu().numpy().view(np.uint64)[0]
