"""Times the resolve kernel of an experimental build (libie_b200_timing.so) on the C4 batch, device-resident."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
if len(sys.argv) > 1:
    ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), sys.argv[1])
from interpolation_engine_b200 import workloads
import torch
eng = ie.Engine(0)
state = workloads.c4_state(); table = eng.pack(state)
n = 1 << 20
sh = workloads.c4_templates(n)
dev = torch.device('cuda', 0)
d_t = torch.from_numpy(sh.bytes).to(dev); d_o = torch.from_numpy(sh.offs.view(np.int64)).to(dev)
cap = int(sh.bytes.nbytes * 2.2) + (1 << 20)
out = torch.empty(cap, dtype=torch.uint8, device=dev); oo = torch.empty(n, dtype=torch.int64, device=dev)
ol = torch.empty(n, dtype=torch.int32, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev); ax = torch.empty(n, dtype=torch.int32, device=dev)
info = torch.zeros(32, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
s = torch.cuda.Stream(device=dev)
def step():
    eng.resolve_batch_device(table, d_t.data_ptr(), d_o.data_ptr(), n, out.data_ptr(), cap, oo.data_ptr(), ol.data_ptr(), st.data_ptr(), ax.data_ptr(), info.data_ptr(), stream=s.cuda_stream)
for _ in range(5): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(20): step()
e1.record(s); torch.cuda.synchronize()
print(ie.LIB_PATH.split('/')[-1], "ms/step", e0.elapsed_time(e1) / 20)
