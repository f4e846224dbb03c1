"""Times the resolve kernel on batches whose tiles outgrow their tables (dense groups without the density hint; bursts of
long templates inside a batch of short ones), next to the C4 batch as the regression check.  Usage:
    python profiles/exp_overflow.py [libie_b200_<variant>.so]
Prints ms per call, the general-path count and a checksum of (lens, out bytes) so two builds can be compared."""
import os, sys, zlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
if len(sys.argv) > 1:
    ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), sys.argv[1])
from interpolation_engine_b200 import workloads
import torch

eng = ie.Engine(0)
state = workloads.c4_state(); table = eng.pack(state)
dev = torch.device('cuda', 0)
rng = np.random.default_rng(11)


def dense(n):
    return ["".join("w%d {q-%d} " % (k, rng.integers(0, 32768)) for k in range(14)) for _ in range(n)]


def bursts(n):
    out = []
    for i in range(n):
        if (i // 128) % 16 == 3:   # one tile in 16 is made of long templates
            out.append("".join("long text %d {q-%d} and more filler text here; " % (k, rng.integers(0, 32768)) for k in range(12)))
        else:
            out.append("short {q-%d} t" % rng.integers(0, 32768))
    return out


def run(name, arena, limits=None):
    n = arena.n
    d_t = torch.from_numpy(np.array(arena.bytes)).to(dev); d_o = torch.from_numpy(np.array(arena.offs).view(np.int64)).to(dev)
    cap = int(arena.bytes.nbytes * 8) + (1 << 20)
    out = torch.empty(cap, dtype=torch.uint8, device=dev); oo = torch.empty(n, dtype=torch.int64, device=dev)
    ol = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); ax = torch.empty(n, dtype=torch.int32, device=dev)
    info = torch.zeros(64, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    kw = {"limits": limits} if limits else {}
    def step():
        eng.resolve_batch_device(table, d_t.data_ptr(), d_o.data_ptr(), n, out.data_ptr(), cap, oo.data_ptr(), ol.data_ptr(), st.data_ptr(), ax.data_ptr(), info.data_ptr(), stream=s.cuda_stream, **kw)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10): step()
    e1.record(s); torch.cuda.synchronize()
    lens = ol.cpu().numpy(); offs = oo.cpu().numpy(); o = out.cpu().numpy()
    # arena placement is unordered: checksum the strings in template order
    order_sum = 0
    for i in range(0, n, max(1, n // 4096)):
        order_sum = zlib.crc32(o[offs[i]:offs[i] + lens[i]].tobytes(), order_sum)
    inf = info.cpu().numpy().view(np.uint64)
    print("%-22s %-22s n=%d  %.4f ms  n_general=%d  lens_sum=%d crc=%08x" % (ie.LIB_PATH.split('/')[-1], name, n, e0.elapsed_time(e1) / 10, int(inf[2]), int(lens.sum()), order_sum))


run("c4", workloads.c4_templates(1 << 20))
a = ie.Arena.from_strings(dense(200000))
run("dense14 no hint", a, (0, 0, a.bytes.nbytes // a.n, 0, 0))
run("dense14 hint", a, (0, 0, a.bytes.nbytes // a.n, 14, 0))
b = ie.Arena.from_strings(bursts(200000))
run("bursts of long", b, (0, 0, b.bytes.nbytes // b.n, 0, 0))
