"""Sharding of independent work units over GPUs (SURVEY.md §8 e): contiguous ranges, no data-path
collective; the only cross-rank steps are the timing reduction and the host-side gather."""


def shard_range(n_total, rank, world):
    """Contiguous [lo, hi) of rank's units: ceil(n_total / world) per rank, the tail rank takes the rest."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def gather_offsets(shard_out_bytes):
    """Base offset of every shard's result arena in the concatenated result (host prefix over shard totals)."""
    bases, acc = [], 0
    for b in shard_out_bytes:
        bases.append(acc)
        acc += int(b)
    return bases, acc


def reduce_timing(dist, local_ms, local_units):
    """max over ranks of the elapsed time, sum over ranks of the units processed (control plane only)."""
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64)
    u = torch.tensor([float(local_units)], dtype=torch.float64)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
