"""How slow is the general (rescan) path?  C4 with a fraction of the q-values holding an unescaped `{slot-K}` reference
(the reference rescans spliced values, interp.rs:81-83): those templates are punted to ie_resolve_general_kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
from interpolation_engine_b200 import workloads
from tests import oracle_lib

eng = ie.Engine(0)
orc = oracle_lib.load()
base = workloads.c4_state()
keys = ie.Arena(base.keys, base.key_offs).strings()
vals = ie.Arena(base.vals, base.val_offs).strings()
n = 1 << 18
tmpl = workloads.c4_templates(n)
for frac in (0.0, 0.01, 0.1, 0.5):
    rng = np.random.default_rng(7)
    v2 = list(vals)
    for i, k in enumerate(keys):
        if k.startswith(b"q-") and rng.random() < frac:
            v2[i] = vals[i][:20] + b"{slot-%d}" % rng.integers(0, 16384) + vals[i][20:]
    st = ie.PackedInserts(base.keys, base.key_offs, ie.Arena.from_strings(v2).bytes, ie.Arena.from_strings(v2).offs, base.tags)
    table = eng.pack(st)
    eng.resolve_batch(table, tmpl)
    r = eng.resolve_batch(table, tmpl)
    # device-resident timing of the kernels alone
    import torch
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(tmpl.bytes).to(dev); d_o = torch.from_numpy(tmpl.offs.view(np.int64)).to(dev)
    cap = int(tmpl.bytes.nbytes * 3) + (1 << 20)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.int64, device=dev)] + [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3)] + [torch.zeros(64, dtype=torch.uint8, device=dev)]
    sdev = torch.cuda.Stream(device=dev)
    dev_ms = {}
    for rounds in (0, 2):
        def step():
            eng.resolve_batch_device(table, d_t.data_ptr(), d_o.data_ptr(), n, bufs[0].data_ptr(), cap, bufs[1].data_ptr(), bufs[2].data_ptr(), bufs[3].data_ptr(), bufs[4].data_ptr(), bufs[5].data_ptr(), stream=sdev.cuda_stream, limits=(0, 0, 0, 0, rounds))
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sdev)
        for _ in range(5): step()
        e1.record(sdev); torch.cuda.synchronize()
        dev_ms[rounds] = e0.elapsed_time(e1) / 5
    out, offs, status, aux = orc.build_table(st).resolve_batch(tmpl.bytes, tmpl.offs, threads=16)
    ok = np.array_equal(r.status, status) and orc.first_mismatch(r.out, r.offs, out, offs[:-1], (offs[1:] - offs[:-1]).astype(np.uint32)) is None
    print(f"frac {frac:4.2f}: templates left to the general path (host API, 2 rounds) {r.n_general:7d} of {n}, host-API kernel_ms {r.kernel_ms:8.3f}, device-resident ms/step: no rounds {dev_ms[0]:7.3f}, 2 rounds {dev_ms[2]:7.3f}, parity {'ok' if ok else 'MISMATCH'}")
