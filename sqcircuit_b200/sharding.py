"""Sharding of a sweep over ranks (one process per GPU).

Sweep points are independent diagonalisations, so the only communication of the path is the final
gather of the per-point results (`gather_to_root`).  Used by bench.py and covered on CPU with
gloo in tests/test_multirank.py."""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_points, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_points for `rank` of `world`."""
    base, rem = divmod(n_points, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(values, rank, world):
    lo, hi = shard_bounds(len(values), rank, world)
    return np.asarray(values)[lo:hi]


def gather_to_root(local: torch.Tensor, n_points: int, root=0):
    """Gather per-point rows [n_local, ...] from every rank into [n_points, ...] on `root`
    (returns None elsewhere).  Shards may differ in size by one row: padded to the max."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_points, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == root else None
    dist.gather(pad, bufs, dst=root)
    if rank != root:
        return None
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
