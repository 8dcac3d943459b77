"""Multi-GPU sharding of the hot path: blob-parallel replicas.

Every GPU holds the full (small) tables and processes its own blobs; there is no data-path
collective.  torch.distributed is used for the rendezvous, the barrier around the timed region and
the max-over-ranks of the measured time only."""
from __future__ import annotations


def blob_range(rank: int, world: int, blobs_per_gpu: int) -> range:
    """Global blob indices owned by `rank` for one step (weak scaling: fixed work per GPU)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return range(rank * blobs_per_gpu, (rank + 1) * blobs_per_gpu)


def blob_seed(blob_index: int) -> int:
    """Seed of the synthetic polynomial of a blob (go_kzg_b200.synth)."""
    return 0xB2000000 + blob_index


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Max of a per-rank scalar over the default process group (identity without a group)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(blobs_per_gpu: int, world: int, steps: int, seconds_max_over_ranks: float) -> float:
    """blobs/s of the whole job: all ranks' blobs over the slowest rank's time."""
    return world * blobs_per_gpu * steps / seconds_max_over_ranks
