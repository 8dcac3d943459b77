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


def merge_sharded(world: int, k2: int) -> bool:
    """Whether the last log2(world) forward stages are sharded too (b200_fk20_multi_finish_merge_part_dev): every rank needs
    at least one position per block."""
    return block_sharded_transforms(world, k2) and world * world <= k2


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
        if merge_sharded(world, k2):
            d_mine = torch.zeros((k2 // world, 18), dtype=torch.int64, device="cuda")
            _check(L.b200_fk20_multi_finish_merge_part_dev(fk.h, blocks.data_ptr(), rank, world, d_mine.data_ptr(), sp), "FK20 multi finish (merge part)")
            parts2 = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(parts2, d_mine)
            _check(L.b200_fk20_multi_finish_assemble_dev(fk.h, parts2.data_ptr(), world, 1, d_out.data_ptr(), sp), "FK20 multi finish (assemble)")
        else:
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


def measure_da_using_fk20_multi_sharded(scale: int = 21, chunk_len: int = 16, reps: int = 2, dist=None, secret: int = 1927409816240961209460912649124):
    """Config 5 of BASELINE.json as a measurement: DAUsingFK20Multi over n = 2^(scale-1) coefficients with the chunk
    offsets sharded over the ranks of the default process group (one rank: the plain single-GPU pipeline through the
    same building blocks).  Every rank builds ONLY the xExtFFT files and window tables of its own offsets.
    Device-resident, CUDA events on the current stream, max over ranks, best of `reps` after one warm-up.
    Returns a dict (identical on every rank)."""
    import time
    import torch
    from .synth import random_fr_limbs
    L = kzg.lib()
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    n = 1 << (scale - 1)
    k2 = 2 * n // chunk_len
    t0 = time.perf_counter()
    setup = kzg.generate_testing_setup_g1(secret, 1 << scale)
    fs = kzg.FFTSettings(scale)
    ks = kzg.KZGSettings(fs, setup)
    del setup
    mine = offset_range(rank, world, chunk_len)
    fk = kzg.FK20MultiSettings(ks, 1 << scale, chunk_len, offsets=mine)
    setup_s = time.perf_counter() - t0
    poly = random_fr_limbs(n, 5)
    d_poly = torch.from_numpy(poly.view(np.int64)).cuda()
    sp = _stream_ptr(torch)
    d_part = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    parts = torch.zeros((world, k2, 18), dtype=torch.int64, device="cuda")
    d_sum = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    d_out = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    d_block = torch.zeros((max(1, k2 // world), 18), dtype=torch.int64, device="cuda")
    blocks = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")
    sharded_transforms = block_sharded_transforms(world, k2)
    sharded_merge = merge_sharded(world, k2)
    d_mine = torch.zeros((max(1, k2 // world), 18), dtype=torch.int64, device="cuda")
    parts2 = torch.zeros((k2, 18), dtype=torch.int64, device="cuda")

    def step(ev):
        ev[0].record()
        _check(L.b200_fk20_multi_partial_dev(fk.h, d_poly.data_ptr(), n, mine.start, mine.stop, d_part.data_ptr(), sp), "FK20 multi partial")
        ev[1].record()
        if world > 1:
            dist.all_gather_into_tensor(parts, d_part)
            _check(L.b200_g1_sum_dev(parts.data_ptr(), world, k2, d_sum.data_ptr(), sp), "G1 sum")
            src = d_sum
        else:
            src = d_part
        ev[2].record()
        if sharded_transforms:
            _check(L.b200_fk20_multi_finish_local_dev(fk.h, src.data_ptr(), rank, world, d_block.data_ptr(), sp), "finish (local)")
            dist.all_gather_into_tensor(blocks, d_block)
            if sharded_merge:
                _check(L.b200_fk20_multi_finish_merge_part_dev(fk.h, blocks.data_ptr(), rank, world, d_mine.data_ptr(), sp), "finish (merge part)")
                dist.all_gather_into_tensor(parts2, d_mine)
                _check(L.b200_fk20_multi_finish_assemble_dev(fk.h, parts2.data_ptr(), world, 1, d_out.data_ptr(), sp), "finish (assemble)")
            else:
                _check(L.b200_fk20_multi_finish_merge_dev(fk.h, blocks.data_ptr(), world, 1, d_out.data_ptr(), sp), "finish (merge)")
        else:
            _check(L.b200_fk20_multi_finish_dev(fk.h, src.data_ptr(), 1, d_out.data_ptr(), sp), "finish")
        ev[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    step(ev)
    barrier()
    best = None
    for _ in range(reps):
        barrier()
        step(ev)
        barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])],
                         device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t = t.tolist()
        if best is None or t[3] < best[3]:
            best = t
    chk = d_out[:: max(1, k2 // 64)].contiguous()
    same = True
    if world > 1:
        ref = chk.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([int(torch.equal(ref, chk))], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    # one position against the closed form (fk20_multi_test.go:60-91): proof = q(s) G, p = q (X^l - x^l) + r, computed on the host
    pos = 12345 % k2
    p_int = kzg.fr_to_ints(poly)
    R = kzg.R_MOD
    bits = k2.bit_length() - 1
    brp = int(format(pos, "0%db" % bits)[::-1], 2)
    rbuf = np.zeros((1, 4), dtype=np.uint64)
    L.b200_fr_root_of_unity(scale, rbuf.ctypes.data)           # root of the full 2n domain: x = w_2n^brp(pos) (fk20_multi_test.go:60-64)
    root = kzg.fr_to_ints(rbuf)[0]
    xl = pow(pow(root, brp, R), chunk_len, R)
    acc = 0
    rem = list(p_int)
    sp_list = [1] * n
    for i in range(1, n):
        sp_list[i] = sp_list[i - 1] * secret % R
    for i in range(n - 1, chunk_len - 1, -1):
        acc = (acc + rem[i] * sp_list[i - chunk_len]) % R
        rem[i - chunk_len] = (rem[i - chunk_len] + rem[i] * xl) % R
    gen = np.zeros(18, dtype=np.uint64)
    L.b200_g1_generator(gen.ctypes.data)
    want = np.zeros(18, dtype=np.uint64)
    kk = kzg.fr_from_ints([acc])
    L.b200_g1_mul(want.ctypes.data, gen.ctypes.data, kk.ctypes.data)
    got = np.ascontiguousarray(d_out[pos].cpu().numpy().view(np.uint64))
    closed = L.b200_g1_equal(got.ctypes.data, want.ctypes.data) == 1
    fk.close(); ks.close(); fs.close()
    return {"workload": "DAUsingFK20Multi n=2^%d chunk=%d -> %d coset proofs, chunk offsets sharded over %d GPU(s)" % (scale - 1, chunk_len, k2, world),
            "n_gpus": world, "ms_total": round(best[3], 2), "ms_partial_hext_fft": round(best[0], 2),
            "ms_exchange_allgather_plus_g1_sum": round(best[1], 2), "ms_g1_transforms": round(best[2], 2),
            "transforms_block_sharded": bool(sharded_transforms), "merge_stages_sharded": bool(sharded_merge), "polys_per_s": round(1e3 / best[3], 4),
            "exchange_bytes_per_rank": int(k2 * 144) if world > 1 else 0, "ranks_agree": same, "closed_form_position_ok": bool(closed),
            "files_per_rank": len(mine), "settings_build_s": round(setup_s, 1)}
