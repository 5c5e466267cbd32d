"""Multi-GPU sharding helpers (host logic only).  Samples and edges are independent units, so each rank owns a contiguous
index range of the counter-generated sample stream and no data-path collective exists; torch.distributed is used only to
agree on the timing (max over ranks) and to sum the per-rank counters."""


def shard_range(n, rank, world):
    """contiguous range [lo, hi) of rank `rank` when n units are split over `world` ranks (sizes differ by at most 1)"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_scalar(x, op, dist=None, device=None):
    """max / sum of a python scalar over the ranks of the default process group (identity when not initialised)"""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def gather_mask(mask, dist=None):
    """concatenate the per-rank uint8 masks in rank order on every rank (the only data that returns to the host side)"""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return mask
    import torch
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, mask)
    import numpy as np
    return np.concatenate(out)
