"""Times the resolve kernel of experimental builds on the C4 batch (device-resident) and checks their results.

    python profiles/exp_variant.py [libie_b200_x.so ...]      # each in its own process; the default build goes first

Variants are built with `make -C interpolation_engine_b200/csrc NAME=x DEFS=-D...`.  Every run prints the kernel time
and a digest of (status, length, bytes) of all results in template order, so a variant that changes any output byte
shows up as a different digest.
"""
import os
import subprocess
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(libname):
    import torch

    import interpolation_engine_b200 as ie
    from interpolation_engine_b200 import workloads
    if libname:
        ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), libname)
    eng = ie.Engine(0)
    state = workloads.c4_state()
    table = eng.pack(state)
    n = 1 << 20
    sh = workloads.c4_templates(n)
    dev = torch.device('cuda', 0)
    d_t = torch.from_numpy(sh.bytes).to(dev)
    d_o = torch.from_numpy(sh.offs.view(np.int64)).to(dev)
    cap = int(sh.bytes.nbytes * 2.2) + (1 << 20)
    out = torch.empty(cap, dtype=torch.uint8, device=dev)
    oo = torch.empty(n, dtype=torch.int64, device=dev)
    ol = torch.empty(n, dtype=torch.int32, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    ax = torch.empty(n, dtype=torch.int32, device=dev)
    info = torch.zeros(64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(device=dev)

    def step():
        eng.resolve_batch_device(table, d_t.data_ptr(), d_o.data_ptr(), n, out.data_ptr(), cap, oo.data_ptr(), ol.data_ptr(), st.data_ptr(),
                                 ax.data_ptr(), info.data_ptr(), stream=s.cuda_stream)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    ev[0].record(s)
    for k in range(20):
        step()
        ev[k + 1].record(s)
    torch.cuda.synchronize()
    per = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(20))
    # digest of the results in template order (the arena layout itself is free: tiles land in completion order)
    offs = oo.cpu().numpy().astype(np.int64)
    lens = ol.cpu().numpy().astype(np.int64)
    arena = out.cpu().numpy()
    crc = 0
    for lo in range(0, n, 1 << 16):
        o, l = offs[lo:lo + (1 << 16)], lens[lo:lo + (1 << 16)]
        idx = np.repeat(o - np.concatenate(([0], np.cumsum(l)[:-1])), l) + np.arange(int(l.sum()))
        crc = zlib.crc32(arena[idx].tobytes(), crc)
    crc = zlib.crc32(ol.cpu().numpy().tobytes(), crc)
    crc = zlib.crc32(st.cpu().numpy().tobytes(), crc)
    print(f"{(libname or 'libie_b200.so'):28s} ms/step mean {ev[0].elapsed_time(ev[20]) / 20:.4f} min {per[0]:.4f} median {per[10]:.4f}  digest {crc:08x}",
          flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--one':
        run_one(sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != '-' else None)
    else:
        for lib in ['-'] + sys.argv[1:]:
            subprocess.call([sys.executable, os.path.abspath(__file__), '--one', lib])
