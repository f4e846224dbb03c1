"""Per-phase cycle shares of the escape kernel (build: make NAME=timing DEFS=-DIE_PHASE_TIMING)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), 'libie_b200_timing.so')
from interpolation_engine_b200 import workloads
eng = ie.Engine(0)
tmpl = workloads.c4_templates(1 << 20)
lib = ctypes.CDLL(ie.LIB_PATH)
buf = (ctypes.c_ulonglong * 16)()
names = ['loop sync', 'acquire', 'pass1+publish', 'lookback', 'pass2', '-', '-', '-', '-', '-', '-']
for mode in (0, 1):
    eng.escape_batch(tmpl, mode)
    lib.ie_debug_escape_cycles(buf, 1)
    eng.escape_batch(tmpl, mode)
    lib.ie_debug_escape_cycles(buf, 1)
    tiles = tmpl.bytes.nbytes / 8192 / 8
    tot = sum(buf[k] for k in range(11))
    print('mode', mode)
    for k, nm in enumerate(names):
        print(f"  {nm:18s} {buf[k]/tiles:9.0f} cyc/tile {100*buf[k]/tot:5.1f}%")
    print('  total/tile', tot / tiles)
    print('  lookbacks', buf[12], 'rounds/lookback', buf[13] / max(buf[12], 1), 'spins/lookback', buf[14] / max(buf[12], 1))

tr = (ctypes.c_ulonglong * (4 * 32768))()
eng.escape_batch(tmpl, 0)
lib.ie_debug_escape_trace(tr)
a = np.frombuffer(tr, dtype=np.uint64).reshape(-1, 4)[:int(tiles)].astype(np.int64)
a -= a[:, 0].min()
print('tile: acquire publish lookback_done end (ns)')
for k in list(range(0, 12)) + list(range(1000, 1006)):
    print(k, a[k].tolist())
d = a[:, 1] - a[:, 0]
print('pre-publish ns: mean', d.mean(), 'p50', np.percentile(d, 50), 'p99', np.percentile(d, 99), 'max', d.max())
w = a[:, 2] - a[:, 1]
print('lookback ns: mean', w.mean(), 'p50', np.percentile(w, 50), 'p99', np.percentile(w, 99))
print('acquire order monotone frac', float((np.diff(a[:, 0]) >= 0).mean()), 'publish monotone frac', float((np.diff(a[:, 1]) >= 0).mean()))
e = a[:, 3] - a[:, 2]
print('post ns: mean', e.mean())
# who are the stragglers?  running max of predecessors' publish times vs own publish time
pub = a[:, 1]
rm = np.maximum.accumulate(pub)
prev_max = np.concatenate([[0], rm[:-1]])
wait = np.maximum(prev_max - pub, 0)
print('ideal wait (prev max publish - own publish) ns: mean', wait.mean(), 'p50', np.percentile(wait, 50), 'p99', np.percentile(wait, 99))
strag = np.where(pub >= rm)[0]  # tiles that set a new running max
print('stragglers (new running max):', len(strag), 'of', len(pub))
ds = d[strag]
print('straggler pre-publish ns: mean', ds.mean(), ' all tiles mean', d.mean())
acq = a[:, 0]
late_acq = acq - np.maximum.accumulate(acq)
print('acquire lateness vs running max acquire (ns): min', late_acq.min(), 'mean', late_acq.mean())
# how far ahead in time is acquire of tile k+1184 vs tile k
per = 148
if len(acq) > 2 * per:
    print('acquire spacing over one residency (ns): mean', (acq[per:] - acq[:-per]).mean())
    print('end->next acquire gap irrelevant; tile duration ns mean', (a[:, 3] - a[:, 0]).mean())
