"""Sharding of independent work units over GPUs (SURVEY.md §8 e): contiguous ranges, no data-path collective.
`shard_range` is the partition ie_resolve_batch_multi applies inside one process (the host gather is ie_shards_gather, in
C); with one process per GPU (bench.py under torchrun) the only cross-rank step is the timing reduction below."""


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of rank's units: ceil(n_total / world) per rank, the tail rank takes the rest."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def reduce_timing(dist, local_ms, local_units):
    """max over ranks of the elapsed time, sum over ranks of the units processed (control plane only)."""
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64)
    u = torch.tensor([float(local_units)], dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
