import sys, os, time, random
sys.path.insert(0, '/root/repo')
import numpy as np
import interpolation_engine_b200 as ie
from tests import oracle_lib, fuzz_campaign as fc
if len(sys.argv) > 3: ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), sys.argv[3])
eng, oracle = ie.Engine(0), oracle_lib.load()
orig_pack_many, orig_resolve = eng.pack_many, eng.resolve_batch
def pm(*a, **k):
    t0 = time.time(); r = orig_pack_many(*a, **k); print('   pack_many %.2fs states=%d' % (time.time() - t0, len(a[0])), flush=True); return r
def rb(table, arena, **k):
    t0 = time.time(); r = orig_resolve(table, arena, **k); print('   resolve_batch %.2fs n=%d kernel_ms=%.2f n_general=%d' % (time.time() - t0, arena.n if hasattr(arena,'n') else len(arena), r.kernel_ms, r.n_general), flush=True); return r
eng.pack_many, eng.resolve_batch = pm, rb
for seed in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    ins, templates = fc.batch(seed) if seed % 3 else fc.big_table_batch(seed)
    bad = []
    t0 = time.time()
    class _Stop(Exception): pass
    real_build = oracle.build_table
    def stop(*a, **k): raise _Stop()
    oracle.build_table = stop
    try: n = fc.many_states(eng, oracle, seed, ins, templates, bad)
    except _Stop: n = -1
    oracle.build_table = real_build
    print('seed %d many_states: %d results in %.2fs, %d bad, %d inserts %d templates' % (seed, n, time.time() - t0, len(bad), len(ins), len(templates)), flush=True)
