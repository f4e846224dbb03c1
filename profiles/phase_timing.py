import ctypes, os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), 'libie_b200_timing.so')
from interpolation_engine_b200 import workloads
eng = ie.Engine(0)
state = workloads.c4_state(); table = eng.pack(state)
tmpl = workloads.c4_templates(1 << 20)
buf = (ctypes.c_ulonglong * 16)()
for it in range(3):
    eng.resolve_batch(table, tmpl)
    eng.lib.ie_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ie._lib.ie_debug_phase_cycles(buf, 1) if False else None
lib = ctypes.CDLL(ie.LIB_PATH)
lib.ie_debug_phase_cycles(buf, 1)
r = eng.resolve_batch(table, tmpl)
lib.ie_debug_phase_cycles(buf, 1)
names = ['-','P0+P1','sync1','P2 events+structure','sync2','P3 lookups','sync3','P4 sizes','scan+emit','claim+offs','P5 passA','P5 passB','-','-']
tiles = (tmpl.n + 127)//128
tot = sum(buf[k] for k in range(14))
for k,nm in enumerate(names): print(f"{nm:14s} {buf[k]/tiles:10.0f} cyc/tile  {100*buf[k]/tot:5.1f}%")
print("total per tile", tot/tiles, "kernel_ms", r.kernel_ms)
