#!/usr/bin/env python
"""Secondary measurements (not the driver's contract line — that is bench.py): the other configs of
BASELINE.json on one B200, one JSON line each, same roofline / cpu_baseline conventions.

    python bench_aux.py [--which c5,escape,c3,c1c2]

The CPU oracle (oracle/, test infrastructure) is loaded here ONLY for the `cpu_baseline` legs and, in c1c2, as the parity
check of the three calls it times; nothing that is measured as "ours" touches it.

  c5      wildcard delete sweep over 10 M keys x 64 pattern sets (runtime.rs:1198-1239, 1633-1647)
  escape  recursive_escape / recursive_unescape over the 1 Mi C4 templates (interp.rs:147-177)
  c3      text_adventure-derived templates over 10 000 cloned states (one launch; and one launch per state)
  c1c2    the hello_world / math task traces, one recursive_interpolate call per task (latency-bound)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bench import measured_peak  # noqa: E402


def device_time_ms(torch, stream, fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_c5(eng, ie, workloads, torch, dev, orc):
    keys = workloads.c5_keys()
    sets = workloads.c5_pattern_sets()
    n = keys.n
    d_k = torch.from_numpy(keys.bytes).to(dev)
    d_o = torch.from_numpy(keys.offs.view(np.int64)).to(dev)
    d_m = torch.empty((n + 31) // 32 + 1, dtype=torch.int32, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    arenas = [ie.Arena.from_strings(p) for p in sets]
    state = {"k": 0}

    def step():
        pa = arenas[state["k"] % len(arenas)]
        eng.glob_sweep_device(d_k.data_ptr(), d_o.data_ptr(), n, pa, state["k"] & 1, d_m.data_ptr(), d_n.data_ptr(), stream=stream.cuda_stream)
        state["k"] += 1
    ms = device_time_ms(torch, stream, step, 2 * len(arenas))
    alg = keys.bytes.nbytes + (n + 1) * 8 + (n + 7) // 8
    peak, src = measured_peak()
    # e2e through the host-buffer call, key arena and offsets in pinned host memory (like bench.py's e2e)
    import ctypes

    def pinned(arr):
        p = ctypes.c_void_p()
        eng._check(eng.lib.ie_host_alloc(arr.nbytes, ctypes.byref(p)))
        ctypes.memmove(p.value, arr.ctypes.data, arr.nbytes)
        return p
    h_k, h_o = pinned(keys.bytes), pinned(keys.offs)
    mask = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
    nd = ctypes.c_uint64(0)

    def sweep(k):
        pa = arenas[k]
        eng._check(eng.lib.ie_glob_sweep(eng.handle, h_k, h_o, n, pa.bytes.ctypes.data, pa.offs.ctypes.data, pa.n, k & 1, mask.ctypes.data, ctypes.byref(nd)))
    sweep(0)
    t0 = time.perf_counter()
    reps = 8
    for k in range(reps):
        sweep(k)
    e2e_s = (time.perf_counter() - t0) / reps
    eng.lib.ie_host_free(h_k)
    eng.lib.ie_host_free(h_o)
    threads = os.cpu_count() or 1
    sub = 1 << 21
    ksub = ie.Arena(keys.bytes[:int(keys.offs[sub])], keys.offs[:sub + 1])
    t0 = time.perf_counter()
    for k in range(4):
        orc.glob_sweep(ksub.bytes, ksub.offs, arenas[k].bytes, arenas[k].offs, bool(k & 1), threads=threads)
    cpu = 4 * sub / (time.perf_counter() - t0)
    return {
        "metric": "wildcard delete sweep keys/sec (10 M keys, 64 pattern sets of 1-9 wildcards)", "value": n / (ms * 1e-3), "unit": "keys/s",
        "n_gpus": 1, "ms_per_step": ms, "higher_is_better": True, "dtype": "u8", "data": "synthetic", "vs_baseline": None,
        "config": {"workload": "C5: 10M keys persona-<p>/field-<f>, one sweep = one (pattern set, delete|delete_except) pair", "keys": n,
                   "l2": "key arena + offsets 0.31 GB > 126 MB L2"},
        "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak,
                     "traffic": None, "kernel": "ie_glob_kernel", "algorithmic_bytes_per_launch": alg, "peak_source": src + ", of measured"},
        "e2e": {"value": n / e2e_s, "unit": "keys/s", "h2d_bytes_per_step": int(keys.bytes.nbytes + keys.offs.nbytes), "d2h_bytes_per_step": int((n + 31) // 32 * 4 + 8),
                "ms_per_step": e2e_s * 1e3},
        "cpu_baseline": {"value": cpu, "unit": "keys/s", "cores": threads, "kind": "port",
                         "sample": f"first {sub} keys x 4 pattern sets; direct '*' matcher (faster than the reference's regex-compile-per-pair, favours the CPU)"},
        "gpu_launches": 2 * len(arenas),
    }


def bench_escape(eng, ie, workloads, torch, dev, orc):
    tmpl = workloads.c4_templates(1 << 20)
    n = tmpl.n
    d_t = torch.from_numpy(tmpl.bytes).to(dev)
    d_o = torch.from_numpy(tmpl.offs.view(np.int64)).to(dev)
    cap = tmpl.bytes.nbytes * 2 + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_oo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    peak, src = measured_peak()
    res = {}
    for mode, name in ((1, "escape"), (0, "unescape")):
        def step():
            eng._check(eng.lib.ie_escape_batch_device(eng.handle, mode, d_t.data_ptr(), d_o.data_ptr(), n, tmpl.bytes.nbytes, d_out.data_ptr(), cap, d_oo.data_ptr(), stream.cuda_stream))
        ms = device_time_ms(torch, stream, step, 10)
        ob = int(d_oo[-1].item())
        alg = tmpl.bytes.nbytes + ob + 2 * (n + 1) * 8
        res[name] = {"ms": ms, "strings_per_s": n / (ms * 1e-3), "achieved_GBs": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak}
    return {"metric": "recursive_escape / recursive_unescape strings/sec (1 Mi C4 templates)", "value": res["unescape"]["strings_per_s"], "unit": "strings/s",
            "n_gpus": 1, "higher_is_better": True, "dtype": "u8", "data": "synthetic", "vs_baseline": None,
            "config": {"workload": "the C4 template arena through ie_escape_batch_device", "detail": res},
            "roofline": {"bound": "hbm", "achieved": res["unescape"]["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": res["unescape"]["frac"], "traffic": None,
                         "kernel": "ie_escape_kernel", "peak_source": src + ", of measured"}}


def bench_c3(eng, ie, workloads, torch, dev, orc):
    """C3: text_adventure-derived templates x 10 000 cloned states.  Two ways: (a) the states packed into one
    table and resolved in ONE launch (ie_table_pack_many, cross product), timed with and without the host-side
    packing; (b) one table + one launch per state, the way a single interactive run calls the resolver."""
    rng = np.random.default_rng(0xC3)
    arena = ie.Arena.from_strings(workloads.C3_TEMPLATES)
    n_states = 10000
    packs = [ie.PackedInserts.from_dict(workloads.c3_state(s, rng)) for s in range(n_states)]
    t0 = time.perf_counter()
    table = eng.pack_many(packs)
    pack_s = time.perf_counter() - t0
    eng.resolve_batch(table, arena)  # warm-up (buffers grow)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = eng.resolve_batch(table, arena)
    e2e_s = (time.perf_counter() - t0) / reps
    n = n_states * arena.n
    in_b = arena.bytes.nbytes * n_states + table.device_bytes
    alg = in_b + int(r.lens.sum()) + 2 * (n + 1) * 8 + 4 * n
    peak, src = measured_peak()
    # (b) per-state launches on a sample
    sample = 512
    t0 = time.perf_counter()
    for pk in packs[:sample]:
        eng.resolve_batch(eng.pack(pk), arena)
    per_state_s = (time.perf_counter() - t0) * n_states / sample
    t0 = time.perf_counter()
    for pk in packs[:256]:
        orc.build_table(pk).resolve_batch(arena.bytes, arena.offs, threads=1)
    cpu_s = (time.perf_counter() - t0) * n_states / 256
    return {"metric": "C3 text_adventure-derived (state, template) pairs/sec, 10 000 cloned states in one launch", "value": n / (r.kernel_ms * 1e-3),
            "unit": "strings/s", "n_gpus": 1, "ms_per_step": r.kernel_ms, "higher_is_better": True, "dtype": "u8", "data": "synthetic", "vs_baseline": None,
            "config": {"workload": f"C3: {n_states} states x {arena.n} templates = {n} pairs, one packed table set ({table.device_bytes >> 20} MiB), one launch",
                       "general_path_templates": int(r.n_general),
                       "table_build": {"python_concat_plus_call_ms": pack_s * 1e3, "ie_table_pack_many_ms": table.pack_call_s * 1e3,
                                       "device_build_kernels_ms": table.build_ms,
                                       "note": "ie_table_pack_many = upload of the raw packed arrays + build kernels on the device (hash, claim, classify, copy)"}},
            "roofline": {"bound": "hbm", "achieved": alg / (r.kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (r.kernel_ms * 1e-3) / 1e9 / peak,
                         "traffic": None, "kernel": "ie_resolve_fused_kernel", "algorithmic_bytes_per_launch": alg,
                         "note": "template text counted once per state although it is L2-resident after the first", "peak_source": src + ", of measured"},
            "e2e": {"value": n / e2e_s, "unit": "strings/s", "ms_per_step": e2e_s * 1e3, "with_table_build": n / (e2e_s + table.pack_call_s),
                    "h2d_bytes_per_step": int(arena.bytes.nbytes + arena.offs.nbytes), "d2h_bytes_per_step": int(r.lens.sum()) + 20 * n,
                    "per_state_launches": {"value": n / per_state_s, "sample": f"{sample} states, pack + H2D + kernels + D2H each"}},
            "cpu_baseline": {"value": n / cpu_s, "unit": "strings/s", "cores": 1, "kind": "port", "sample": "256 states, 1 thread, table build included"}}


def bench_c1c2(eng, ie, workloads, torch, dev, orc):
    """C1 / C2 (BASELINE configs 0-1): the interactive shape — one task at a time, a handful of templates per call.
    The reference's own CPU path is the right tool here; both sides are timed per recursive_interpolate call."""
    c1 = {"cmd": "print", "text": "Hello, world!", "line": 8}                                   # hello_world: 5 templates, 0 lookups
    c2a = {"cmd": "math", "input": "max(1,2,3)", "output_name": "result", "line": 8}             # math task 1: 7 templates
    c2b = {"cmd": "print", "text": "The result is {result}!\n", "line": 9}                      # math task 2: 5 templates, 1 lookup
    calls = [({}, c1), ({}, c2a), ({"result": 3}, c2b)]
    for ins, task in calls:  # warm-up + parity
        assert eng.call("recursive_interpolate", inserts=ins, value=task) == orc.call("recursive_interpolate", inserts=ins, value=task)
    reps = 200
    t0 = time.perf_counter()
    for _ in range(reps):
        for ins, task in calls:
            eng.call("recursive_interpolate", inserts=ins, value=task)
    gpu_us = (time.perf_counter() - t0) / (reps * len(calls)) * 1e6
    t0 = time.perf_counter()
    for _ in range(reps):
        for ins, task in calls:
            orc.call("recursive_interpolate", inserts=ins, value=task)
    cpu_us = (time.perf_counter() - t0) / (reps * len(calls)) * 1e6
    # the same calls against snapshots that outlive the call (the run loop's pattern: one map, patched in place): no
    # serialisation, packing or upload of the state per call
    sids = {}
    for ins, _ in calls:
        key = json.dumps(ins, sort_keys=True)
        if key not in sids:
            sids[key] = eng.call("snapshot_create", inserts=ins)[1]
    for ins, task in calls:
        assert eng.call("recursive_interpolate", snapshot=sids[json.dumps(ins, sort_keys=True)], value=task) == orc.call("recursive_interpolate", inserts=ins, value=task)
    t0 = time.perf_counter()
    for _ in range(reps):
        for ins, task in calls:
            eng.call("recursive_interpolate", snapshot=sids[json.dumps(ins, sort_keys=True)], value=task)
    snap_us = (time.perf_counter() - t0) / (reps * len(calls)) * 1e6
    # ... and with a state of 65 536 inserts behind them (C4's): the per-call cost must not depend on the state's size
    big = {"slot-%d" % k: k for k in range(65536)}
    big["result"] = 3
    sid_big = eng.call("snapshot_create", inserts=big)[1]
    task = calls[2][1]
    eng.call("recursive_interpolate", snapshot=sid_big, value=task)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.call("recursive_interpolate", snapshot=sid_big, value=task)
    snap_big_us = (time.perf_counter() - t0) / reps * 1e6
    t0 = time.perf_counter()
    eng.call("recursive_interpolate", inserts=big, value=task)
    nosnap_big_us = (time.perf_counter() - t0) * 1e6
    n_t = 17  # templates over the three calls
    return {"metric": "C1/C2 recursive_interpolate calls/sec (hello_world + math traces, one task per call)", "value": 1e6 / gpu_us, "unit": "calls/s",
            "n_gpus": 1, "ms_per_step": gpu_us * 1e-3, "higher_is_better": True, "dtype": "u8", "data": "the two example programs' task objects", "vs_baseline": None,
            "config": {"workload": "C1 + C2: 3 tasks, 17 templates, 1 lookup; JSON in, pack, H2D, two kernels, D2H, JSON out per call",
                       "templates_per_call": n_t / 3},
            "cpu_baseline": {"value": 1e6 / cpu_us, "unit": "calls/s", "cores": 1, "kind": "port", "sample": f"{reps} x 3 calls through the oracle's JSON entry point"},
            "snapshot_calls": {"us_per_call": snap_us, "us_per_call_65536_inserts": snap_big_us, "us_per_call_65536_inserts_without_snapshot": nosnap_big_us,
                               "what": "the same calls with {snapshot: id}: the state stays packed on the device between calls"},
            "note": "latency-bound: a GPU call costs %.0f us (%.0f us against a live snapshot) and %.0f us on one CPU thread; the CPU wins by %.1fx at this batch size, as expected"
                    % (gpu_us, snap_us, cpu_us, min(gpu_us, snap_us) / cpu_us)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c5,escape,c3,c1c2")
    args = ap.parse_args()
    import torch

    import interpolation_engine_b200 as ie
    from interpolation_engine_b200 import workloads
    from tests import oracle_lib
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = ie.Engine(0)
    orc = oracle_lib.load()
    for name in args.which.split(","):
        fn = {"c5": bench_c5, "escape": bench_escape, "c3": bench_c3, "c1c2": bench_c1c2}[name]
        print(json.dumps(fn(eng, ie, workloads, torch, dev, orc)), flush=True)


if __name__ == "__main__":
    main()
