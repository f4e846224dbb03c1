#!/usr/bin/env python
"""Headline benchmark: batched `{key}` interpolation, config C4 of BASELINE.json / SURVEY.md §8(d).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A step = one pass of the hot path over one batch of 1 Mi synthetic templates per GPU against a
65 536-insert state (depth-3 nesting).  `value` is whole-job interpolated strings/s with inputs and
outputs resident in HBM (CUDA events on the launching stream, max over ranks); `e2e` is the same
metric through the host-buffer C-ABI call ie_resolve_batch (pinned host arenas, H2D + kernels + D2H
inside the timed region); `roofline` relates the device time to the algorithmic bytes of
SURVEY.md §8(d); `cpu_baseline` is the oracle (C++ restatement of interp.rs, NOT the Rust binary —
it cannot be built here) on the host cores over a bounded sample.  torch is plumbing only: device
buffers, streams/events and the torch.distributed barrier.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TEMPLATES = 1 << 20
METRIC = "interpolated strings/sec (1 Mi templates x 64k-insert state, depth-3 nesting)"
WORKLOAD = "C4: synthetic 1Mi templates x 65536-insert state, depth-3 nesting, per GPU"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the fused tile resolve kernel on this workload, taken
    from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("ie_resolve_fused_kernel_dram_bytes_per_launch")
    return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons), "samples": len(rows)}


def algorithmic_bytes(tmpl_bytes, out_bytes, n, table_bytes):
    # stream in + stream out + two 64-bit offset arrays + status word + table once
    return tmpl_bytes + out_bytes + 2 * (n + 1) * 8 + 4 * n + table_bytes


def cpu_reference_rate(state, tmpl, sample, threads, repeats=1):
    """Oracle (reference algorithm restated in C++) on `sample` templates with `threads` host threads."""
    from interpolation_engine_b200 import Arena
    from tests import oracle_lib
    orc = oracle_lib.load()
    tab = orc.build_table(state)
    sub = Arena(tmpl.bytes[:int(tmpl.offs[sample])], tmpl.offs[:sample + 1])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        tab.resolve_batch(sub.bytes, sub.offs, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, best


def run_reference(args, rank):
    if rank != 0:
        return
    from interpolation_engine_b200 import workloads
    threads = os.cpu_count() or 1
    state = workloads.c4_state()
    sample = 1 << 18
    tmpl = workloads.c4_templates(sample)
    for _ in range(args.warmup):
        cpu_reference_rate(state, tmpl, sample, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(state, tmpl, sample, threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "strings/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"first {sample} templates of the batch per step"},
        "cpu_baseline": {"value": value, "unit": "strings/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} templates per step, {threads} host threads, C++ restatement of interp.rs (not the Rust binary)"},
        "e2e": {"value": value, "unit": "strings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_other_workload(args, rank, local_rank):
    """BASELINE.json's other configs (wildcard sweep, cloned states, escape / unescape, the two tiny example traces) with
    the contract's line shape: measured by bench_aux.py's functions on ONE GPU (rank 0; these paths shard the same way
    C4 does - independent key / state ranges, no collective - and only C4 is run at 2/4/8 GPUs)."""
    if rank != 0:
        return
    import torch

    import bench_aux
    import interpolation_engine_b200 as ie
    from interpolation_engine_b200 import workloads
    from tests import oracle_lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = ie.Engine(local_rank)
    fn = {"c5": bench_aux.bench_c5, "escape": bench_aux.bench_escape, "c3": bench_aux.bench_c3, "c1c2": bench_aux.bench_c1c2}[args.workload]
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.time()
    line = fn(eng, ie, workloads, torch, dev, oracle_lib.load())
    line["clocks"] = sampler.stop(t0, time.time())
    line.setdefault("steps", args.steps)
    line.setdefault("warmup", max(args.warmup, 3))
    line.setdefault("scaling", "weak")
    line.setdefault("gpu_launches", None)
    line["n_gpus"] = 1
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--templates", type=int, default=N_TEMPLATES, help="templates per GPU (default = the BASELINE config)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5", "c3", "escape", "c1c2"],
                    help="c4 (default) = the headline config; the others are BASELINE.json's remaining configs, one GPU, same line shape")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "c4":
        run_other_workload(args, rank, local_rank)
        return

    import torch
    import torch.distributed as dist

    import interpolation_engine_b200 as ie
    from interpolation_engine_b200 import workloads

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)  # control plane only: no data-path collective
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = ie.Engine(local_rank)
    n = args.templates

    # ---- inputs: this rank's shard of the global template stream; the state is replicated ----
    state = workloads.c4_state()
    table = eng.pack(state)
    shards = [workloads.c4_templates(n, start=(2 * rank + k) * n) for k in range(2)]  # two buffer sets, alternated per step

    def to_dev(a, dtype):
        return torch.from_numpy(np.ascontiguousarray(a).view(dtype)).to(dev)

    sets = []
    info_bytes = 64  # sizeof(ie_batch_info) = 40
    for sh in shards:
        d_t = to_dev(sh.bytes, np.uint8)
        d_o = to_dev(sh.offs.view(np.int64), np.int64)
        cap = int(sh.bytes.nbytes * 2.2) + (1 << 20)
        sets.append({
            "shard": sh, "d_t": d_t, "d_o": d_o, "cap": cap,
            "out": torch.empty(cap, dtype=torch.uint8, device=dev), "out_offs": torch.empty(n, dtype=torch.int64, device=dev),
            "out_lens": torch.empty(n, dtype=torch.int32, device=dev), "status": torch.empty(n, dtype=torch.int32, device=dev),
            "aux": torch.empty(n, dtype=torch.int32, device=dev), "info": torch.zeros(info_bytes, dtype=torch.uint8, device=dev),
        })
    # an explicit non-default stream: the ABI maps a NULL stream to the engine's own stream, and the
    # CUDA events below must sit on the stream the kernels are launched on
    torch.cuda.synchronize()
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0

    def step(k):
        s = sets[k % 2]
        eng.resolve_batch_device(table, s["d_t"].data_ptr(), s["d_o"].data_ptr(), n, s["out"].data_ptr(), s["cap"],
                                 s["out_offs"].data_ptr(), s["out_lens"].data_ptr(), s["status"].data_ptr(), s["aux"].data_ptr(),
                                 s["info"].data_ptr(), stream=stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for k in range(max(args.warmup, 2)):
        step(k)
    torch.cuda.synchronize()
    out_bytes = []
    for s in sets:
        info = s["info"].cpu().numpy()
        ob = int(info[8:16].view(np.uint64)[0])
        assert ob <= s["cap"], "bench out arena too small"
        out_bytes.append(ob)
        s["n_general"] = int(info[16:24].view(np.uint64)[0])

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- timed region: exactly K steps, barrier + synchronize on both sides, CUDA events ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        step(k)
        ev[k + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    total_ms = ev[0].elapsed_time(ev[-1])
    clocks = sampler.stop(t_wall0, t_wall1)
    per_step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]

    # ---- e2e: host-buffer C-ABI call, pinned host arenas, H2D + kernels + D2H per step ----
    e2e = None
    if not args.no_e2e:
        lib = eng.lib
        sh = shards[0]

        def pinned(arr):
            p = ctypes.c_void_p()
            eng._check(lib.ie_host_alloc(arr.nbytes, ctypes.byref(p)))
            ctypes.memmove(p.value, arr.ctypes.data, arr.nbytes)
            return p

        h_t, h_o = pinned(sh.bytes), pinned(sh.offs)
        res = ie._Result()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            eng._check(lib.ie_resolve_batch(eng.handle, table.handle, h_t, h_o, n, None, ctypes.byref(res)))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng._check(lib.ie_resolve_batch(eng.handle, table.handle, h_t, h_o, n, None, ctypes.byref(res)))
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        h2d = int(sh.bytes.nbytes + sh.offs.nbytes)
        lens = np.frombuffer((ctypes.c_char * (n * 4)).from_address(res.out_lens), dtype=np.uint32)
        d2h = int(lens.sum(dtype=np.uint64)) + n * (8 + 4 + 4 + 4) + 32  # result bytes + offs/lens/status/aux + info
        e2e = {"s": e2e_s, "h2d": h2d, "d2h": d2h, "steps": e2e_steps, "kernel_ms": float(res.info.kernel_ms)}
        lib.ie_host_free(h_t)
        lib.ie_host_free(h_o)

    # ---- e2e, single process: ONE host batch of world * n templates sharded over all GPUs by ie_resolve_batch_multi
    # (one host thread + stream set per device inside the call), then the host gather.  Rank 0 drives every GPU of the
    # box while the other ranks wait at the barrier with their devices idle.
    multi = None
    if world > 1 and not args.no_e2e:
        barrier()
        if rank == 0:
            lib = eng.lib
            engines = [eng] + [ie.Engine(d) for d in range(1, world)]
            tables = [table] + [e2.pack(state) for e2 in engines[1:]]
            big = workloads.c4_templates(world * n)
            p_t, p_o = ctypes.c_void_p(), ctypes.c_void_p()
            eng._check(lib.ie_host_alloc(big.bytes.nbytes, ctypes.byref(p_t)))
            eng._check(lib.ie_host_alloc(big.offs.nbytes, ctypes.byref(p_o)))
            ctypes.memmove(p_t.value, big.bytes.ctypes.data, big.bytes.nbytes)
            ctypes.memmove(p_o.value, big.offs.ctypes.data, big.offs.nbytes)
            eh = (ctypes.c_void_p * world)(*[e2.handle for e2 in engines])
            th = (ctypes.c_void_p * world)(*[t2.handle for t2 in tables])
            shards = (ie._ShardResult * world)()

            def call():
                eng._check(lib.ie_resolve_batch_multi(eh, th, world, p_t, p_o, world * n, None, shards))
            for _ in range(2):
                call()
            m_steps = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(m_steps):
                call()
            m_s = (time.perf_counter() - t0) / m_steps
            # the host gather (concatenation in template order), timed by itself
            total = sum(int(np.frombuffer((ctypes.c_char * (int(sh.n) * 4)).from_address(sh.res.out_lens), dtype=np.uint32).sum(dtype=np.uint64)) for sh in shards)
            g_out = np.empty(total, dtype=np.uint8)
            g_offs = np.empty(world * n + 1, dtype=np.uint64)
            g_st = np.empty(world * n, dtype=np.int32)
            g_ax = np.empty(world * n, dtype=np.uint32)
            ob = ctypes.c_uint64(0)
            t0 = time.perf_counter()
            eng._check(lib.ie_shards_gather(shards, world, g_out.ctypes.data, total, g_offs.ctypes.data, g_st.ctypes.data, g_ax.ctypes.data, ctypes.byref(ob)))
            gather_s = time.perf_counter() - t0
            multi = {"value": world * n / m_s, "unit": "strings/s", "ms_per_step": m_s * 1e3, "host_gather_ms": gather_s * 1e3,
                     "with_host_gather": world * n / (m_s + gather_s), "steps": m_steps,
                     "what": "ie_resolve_batch_multi: one process, one host batch of %d templates, contiguous shards over %d GPUs, one host thread per device; "
                             "ie_shards_gather concatenates %d result bytes in template order" % (world * n, world, total)}
            for t2 in tables[1:]:
                t2.free()
            for e2 in engines[1:]:
                e2.close()
            lib.ie_host_free(p_t)
            lib.ie_host_free(p_o)
        barrier()

    # ---- reduce over ranks: max time, summed work ----
    from interpolation_engine_b200 import sharding
    total_ms_max, _ = sharding.reduce_timing(dist if world > 1 else None, total_ms, n * args.steps)
    e2e_s_max, _ = sharding.reduce_timing(dist if world > 1 else None, e2e["s"] if e2e else 0.0, n)

    if rank == 0:
        ms_per_step = total_ms_max / args.steps
        value = world * n * args.steps / (total_ms_max * 1e-3)
        peak, peak_src = measured_peak()
        tb = [int(s["shard"].bytes.nbytes) for s in sets]
        alg = [algorithmic_bytes(tb[k], out_bytes[k], n, table.device_bytes) for k in range(2)]
        alg_mean = sum(alg[k % 2] for k in range(args.steps)) / args.steps
        step_ms_rank0 = total_ms / args.steps
        achieved = alg_mean / (step_ms_rank0 * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "strings/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 2),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "templates_per_gpu": n, "inserts": int(state.n), "sharding": f"independent template shards x{world}, table replicated, no collective",
                       "mean_template_bytes": tb[0] / n, "mean_output_bytes": out_bytes[0] / n, "general_path_templates": sets[0]["n_general"],
                       "l2": "two input/output buffer sets alternated per step; per-step working set %.2f GB > 126 MB L2" % (alg[0] / 1e9),
                       "timing": "CUDA events on the launching stream, max over ranks; per-step min/median ms = %.3f/%.3f" % (min(per_step_ms), sorted(per_step_ms)[len(per_step_ms) // 2])},
            "clocks": clocks,
            "gpu_launches": 3 * args.steps,  # per rank and step: ie_resolve_fused_kernel + the two tiers of ie_resolve_general_kernel (profiles/r02_final_launches.csv)
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                         "kernel": "ie_resolve_fused_kernel (+ the two general-tier kernels, empty on this workload, and two small memsets in the same step)",
                         "algorithmic_bytes_per_launch": alg_mean, "peak_source": peak_src + ", of measured"},
        }
        if e2e:
            line["e2e"] = {"value": world * n / e2e_s_max, "unit": "strings/s", "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                           "ms_per_step": e2e_s_max * 1e3, "kernel_ms_inside": e2e["kernel_ms"], "timing": "wall clock around ie_resolve_batch (synchronous), pinned host arenas"}
            if multi:
                line["e2e"]["single_process"] = multi
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample = n
            rate, secs = cpu_reference_rate(state, shards[0], sample, threads, repeats=5)
            line["cpu_baseline"] = {"value": rate, "unit": "strings/s", "cores": threads, "kind": "port",
                                    "sample": f"all {sample} templates of the batch, best of 5 ({secs:.2f} s each = {secs * threads:.0f} core-seconds), {threads} host threads; "
                                              "C++ restatement of the reference algorithm, not the Rust binary"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
