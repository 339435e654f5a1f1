"""Multi-GPU plumbing of the path: images are sharded by rank, every rank decodes,
samples and scores its own images with no communication, and the reduced score
tables are exchanged with ONE all-gather (NCCL over NVLink on the GPUs; the same
code runs over gloo in the CPU tests).  Replaces the reference's single-process
``nn.DataParallel`` scatter / gather (OSIE/test.py:94-95)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of rank; sizes differ by at most one."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def padded_shard(n_total: int, world: int) -> int:
    return (n_total + world - 1) // world


def allgather_tables(table: torch.Tensor, n_total: int, image_dim: int = -2, group=None) -> torch.Tensor:
    """table: this rank's [..., n_local, C] slice along `image_dim` (shard_range order).
    Returns the full [..., n_total, C] table on every rank with a single all_gather_into_tensor
    of equal-sized (padded) blocks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return table
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = table.movedim(image_dim, 0).contiguous()
    pad = padded_shard(n_total, world)
    block = t.new_zeros((pad,) + tuple(t.shape[1:]))
    block[:t.shape[0]] = t
    out = t.new_empty((world * pad,) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, block, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        pieces.append(out[r * pad:r * pad + (hi - lo)])
    return torch.cat(pieces, 0).movedim(0, image_dim)
