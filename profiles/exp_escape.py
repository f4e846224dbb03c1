"""Times ie_escape_batch_device (escape / unescape of the C4 template arena) for experimental builds and checks the output digest.
    python profiles/exp_escape.py [libie_b200_x.so ...]"""
import os, subprocess, sys, zlib
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
    tmpl = workloads.c4_templates(1 << 20)
    n = tmpl.n
    dev = torch.device('cuda', 0)
    d_t = torch.from_numpy(tmpl.bytes).to(dev)
    d_o = torch.from_numpy(tmpl.offs.view(np.int64)).to(dev)
    cap = tmpl.bytes.nbytes * 2 + 64
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_oo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    s = torch.cuda.Stream(device=dev)
    out = []
    for mode, name in ((1, 'escape'), (0, 'unescape')):
        def step():
            eng._check(eng.lib.ie_escape_batch_device(eng.handle, mode, d_t.data_ptr(), d_o.data_ptr(), n, tmpl.bytes.nbytes, d_out.data_ptr(), cap, d_oo.data_ptr(), s.cuda_stream))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record(s)
        for k in range(10):
            step(); ev[k + 1].record(s)
        torch.cuda.synchronize()
        per = sorted(ev[k].elapsed_time(ev[k + 1]) for k in range(10))
        ob = int(d_oo[-1].item())
        crc = zlib.crc32(d_out[:ob].cpu().numpy().tobytes(), zlib.crc32(d_oo.cpu().numpy().tobytes()))
        out.append(f"{name} median {per[5]:.4f} min {per[0]:.4f} digest {crc:08x}")
    print(f"{(libname or 'libie_b200.so'):26s} " + "  ".join(out), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--one':
        run_one(sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != '-' else None)
    else:
        for lib in ['-'] + sys.argv[1:]:
            subprocess.call([sys.executable, os.path.abspath(__file__), '--one', lib])
