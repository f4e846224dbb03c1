"""N-rank host<->device copy probe: what the box's host links and host memory can move when every GPU copies at once.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 profiles/copy_probe.py

Every rank (one per GPU, like bench.py) moves the C4 batch's traffic through page-locked buffers, all ranks starting
together behind a barrier: 191 MB host->device, 305 MB device->host, and both at once (the shape of the e2e pipeline);
the host->device side once from ordinary pinned memory and once from write-combined pinned memory.  Rank 0 prints one
JSON line with per-rank and aggregate GB/s.  This separates "the machine cannot move more" from "our pipeline leaves
bandwidth unused" in the end-to-end scaling numbers.
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

H2D_BYTES = 191 << 20
D2H_BYTES = 305 << 20
REPS = 8


def single_process(n_dev):
    """The same traffic issued from ONE process to n_dev GPUs (asynchronous copies on per-device streams): separates what
    eight processes with eight sets of pinned buffers cost from what the box can move."""
    import torch

    import interpolation_engine_b200 as ie
    lib = ie.load_library()
    devs = []
    for d in range(n_dev):
        torch.cuda.set_device(d)
        dev = torch.device("cuda", d)
        bufs = {}
        for name, nbytes in (("in", H2D_BYTES), ("out", D2H_BYTES)):
            p = ctypes.c_void_p()
            assert lib.ie_host_alloc(nbytes, ctypes.byref(p)) == 0
            ctypes.memset(p.value, 1, nbytes)
            bufs["h_" + name] = torch.from_numpy(np.frombuffer((ctypes.c_char * nbytes).from_address(p.value), dtype=np.uint8))
        bufs["d_in"] = torch.empty(H2D_BYTES, dtype=torch.uint8, device=dev)
        bufs["d_out"] = torch.ones(D2H_BYTES, dtype=torch.uint8, device=dev)
        bufs["s_in"], bufs["s_out"] = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        devs.append(bufs)

    def issue(which):
        for b in devs:
            if which in ("h2d", "duplex"):
                with torch.cuda.stream(b["s_in"]):
                    b["d_in"].copy_(b["h_in"], non_blocking=True)
            if which in ("d2h", "duplex"):
                with torch.cuda.stream(b["s_out"]):
                    b["h_out"].copy_(b["d_out"], non_blocking=True)

    def sync():
        for d in range(n_dev):
            torch.cuda.synchronize(d)
    out = {}
    for which, nbytes in (("h2d", H2D_BYTES), ("d2h", D2H_BYTES), ("duplex", H2D_BYTES + D2H_BYTES)):
        issue(which)
        sync()
        t0 = time.perf_counter()
        for _ in range(REPS):
            issue(which)
        sync()
        dt = (time.perf_counter() - t0) / REPS
        out[which + "_aggregate_GBs"] = n_dev * nbytes / dt / 1e9
        out[which + "_ms"] = dt * 1e3
    print(json.dumps({"probe": "host<->device copies from ONE process", "n_gpus": n_dev, "reps": REPS, **out}), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--single":
        single_process(int(sys.argv[2]))
        return
    import torch
    import torch.distributed as dist

    import interpolation_engine_b200 as ie
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = ie.load_library()

    def host_alloc(nbytes, wc=False):
        p = ctypes.c_void_p()
        st = (lib.ie_host_alloc_wc if wc else lib.ie_host_alloc)(nbytes, ctypes.byref(p))
        assert st == 0
        ctypes.memset(p.value, 1, nbytes)
        return p

    h_in, h_in_wc, h_out = host_alloc(H2D_BYTES), host_alloc(H2D_BYTES, wc=True), host_alloc(D2H_BYTES)
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device=dev)
    d_out = torch.ones(D2H_BYTES, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    # the copies go through torch tensors that alias the page-locked blocks (torch recognises page-locked memory by
    # cudaPointerGetAttributes, so copy_(non_blocking=True) is a plain cudaMemcpyAsync on the current stream)
    def as_tensor(p, nbytes):
        arr = np.frombuffer((ctypes.c_char * nbytes).from_address(p.value), dtype=np.uint8)
        return torch.from_numpy(arr)

    t_in, t_in_wc, t_out = as_tensor(h_in, H2D_BYTES), as_tensor(h_in_wc, H2D_BYTES), as_tensor(h_out, D2H_BYTES)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        fn()  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(REPS):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / REPS
        barrier()
        return dt

    def h2d(src):
        def f():
            with torch.cuda.stream(s_in):
                d_in.copy_(src, non_blocking=True)
        return f

    def d2h():
        with torch.cuda.stream(s_out):
            t_out.copy_(d_out, non_blocking=True)

    def duplex(src):
        def f():
            h2d(src)()
            d2h()
        return f

    assert t_in.is_pinned() and t_in_wc.is_pinned() and t_out.is_pinned()
    res = {
        "h2d_GBs": H2D_BYTES / timed(h2d(t_in)) / 1e9,
        "h2d_wc_GBs": H2D_BYTES / timed(h2d(t_in_wc)) / 1e9,
        "d2h_GBs": D2H_BYTES / timed(d2h) / 1e9,
    }
    dt = timed(duplex(t_in))
    res["duplex_GBs"] = (H2D_BYTES + D2H_BYTES) / dt / 1e9
    res["duplex_ms"] = dt * 1e3
    dt = timed(duplex(t_in_wc))
    res["duplex_wc_GBs"] = (H2D_BYTES + D2H_BYTES) / dt / 1e9
    res["duplex_wc_ms"] = dt * 1e3
    allres = [res]
    if world > 1:
        allres = [None] * world
        dist.all_gather_object(allres, res)
    if rank == 0:
        agg = {k: sum(r[k] for r in allres) for k in res if k.endswith("GBs")}
        line = {"probe": "host<->device copies, all ranks at once", "n_gpus": world, "h2d_bytes": H2D_BYTES, "d2h_bytes": D2H_BYTES, "reps": REPS,
                "aggregate": agg, "per_rank_min": {k: min(r[k] for r in allres) for k in res}, "per_rank_max": {k: max(r[k] for r in allres) for k in res},
                "host_cpus": os.cpu_count(),
                "floor_ms_per_1Mi_batch_per_rank": max(r["duplex_ms"] for r in allres)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
