"""Sharding of a batch of independent bases across ranks (one process per GPU).

A factorization is a sequential pivot chain and never spans GPUs; batches shard by basis
index with no data-path collective.  The only communication is the
host-side gather of results and the max-over-ranks of the device time."""


def shard_range(n_total, rank, world):
    """Contiguous index range [lo, hi) of rank `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value, group=None, device=None):
    """max of a host float over all ranks (identity without torch.distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


def gather_rows(local_rows, n_total, group=None):
    """Gather per-rank result blocks (numpy [n_local, m]) on every rank in index order."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_rows
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, local_rows, group=group)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == n_total
    return out
