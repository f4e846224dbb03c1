"""General path on LONG texts (an LLM answer with a stray brace): ms per template against its length."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
import torch
eng = ie.Engine(0)
table = eng.pack({"name": "Ada", "n": 3})
dev = torch.device("cuda", 0)
for length, count in ((1024, 256), (1700, 256), (4096, 256), (16384, 64), (16384, 2048)):
    body = ("The guard looks at {name}. " * (length // 27 + 1))[:length]
    templates = [body + " stray } brace %d" % i for i in range(count)]
    arena = ie.Arena.from_strings(templates)
    n = arena.n
    d_t = torch.from_numpy(np.array(arena.bytes)).to(dev); d_o = torch.from_numpy(np.array(arena.offs).view(np.int64)).to(dev)
    cap = arena.bytes.nbytes * 2 + (1 << 20)
    out = torch.empty(cap, dtype=torch.uint8, device=dev); oo = torch.empty(n, dtype=torch.int64, device=dev)
    ol = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); ax = torch.empty(n, dtype=torch.int32, device=dev)
    info = torch.zeros(64, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    def step():
        eng.resolve_batch_device(table, d_t.data_ptr(), d_o.data_ptr(), n, out.data_ptr(), cap, oo.data_ptr(), ol.data_ptr(), st.data_ptr(), ax.data_ptr(), info.data_ptr(), stream=s.cuda_stream, limits=(4096, 1 << 16, 0, 0, 0))
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s); step(); step(); e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    inf = info.cpu().numpy().view(np.uint64)
    print("%6d B x %5d templates: %8.3f ms per call, n_general=%d, status %s, %.1f ns per byte of one template" % (length, count, ms, int(inf[2]), np.bincount(st.cpu().numpy() & 0xFF), ms * 1e6 / length), flush=True)
