"""N>1 host logic on CPU: world_size-2 gloo group, shard ranges, max-over-ranks timing, gather bases."""
import os
import socket

import torch.distributed as dist
import torch.multiprocessing as mp

from interpolation_engine_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(1_000_003, rank, world)
    ms, units = sharding.reduce_timing(dist, 10.0 + rank, hi - lo)
    dist.barrier()
    out.put((rank, lo, hi, ms, units))
    dist.destroy_process_group()


def test_world_size_2_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, ms0, u0), (r1, lo1, hi1, ms1, u1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 500_002, 500_002, 1_000_003)
    assert ms0 == ms1 == 11.0 and u0 == u1 == 1_000_003.0


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1 << 20, 10_000_000):
        for world in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
