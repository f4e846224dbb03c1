"""Per-leg wall-clock of the first batches of tests/fuzz_campaign.py (which leg is slow?).  python profiles/diag_fuzz_legs.py <first_seed> <count>"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import interpolation_engine_b200 as ie
from tests import oracle_lib, fuzz_campaign as fc
eng, oracle = ie.Engine(0), oracle_lib.load()
for seed in range(int(sys.argv[1]), int(sys.argv[1]) + int(sys.argv[2])):
    ins, templates = fc.batch(seed) if seed % 3 else fc.big_table_batch(seed)
    packed = ie.PackedInserts.from_dict(ins)
    print("seed %d: %d inserts, %d templates, %d bytes, longest %d" % (seed, len(ins), len(templates), sum(len(t) for t in templates), max(len(t) for t in templates)), flush=True)
    t0 = time.time(); table = eng.pack(packed, hhmm="12:34", hhmmss="12:34:56"); t_pack = time.time() - t0
    arena = ie.Arena.from_strings(templates)
    t0 = time.time(); out, offs, status, aux = oracle.build_table(packed).resolve_batch(arena.bytes, arena.offs, threads=8, hhmm="12:34", hhmmss="12:34:56"); t_or = time.time() - t0
    print("  oracle %.2f s" % t_or, flush=True)
    t0 = time.time(); got = eng.resolve_batch(table, arena, limits=(4096, 1 << 16)); t_host = time.time() - t0
    print("  host call %.2f s kernel_ms %.1f general %d" % (t_host, got.kernel_ms, got.n_general), flush=True)
    n, nb = arena.n, arena.bytes.nbytes
    cap = int((offs[1:] - offs[:-1]).sum()) * 2 + (1 << 20)
    t_dev = []
    for rounds in (0, 1, 2, 3):
        d_t = eng.alloc(nb + 64).upload(arena.bytes); d_o = eng.alloc((n + 1) * 8).upload(arena.offs)
        bufs = (eng.alloc(cap + 64), eng.alloc(n * 8), eng.alloc(n * 4), eng.alloc(n * 4), eng.alloc(n * 4), eng.alloc(64))
        eng.sync(); t0 = time.time()
        eng.resolve_batch_device(table, d_t.ptr, d_o.ptr, n, bufs[0].ptr, cap, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr, bufs[4].ptr, bufs[5].ptr, limits=(4096, 1 << 16, 0, 0, rounds))
        eng.sync(); t_dev.append(time.time() - t0)
        print("  device call rounds=%d %.3f s" % (rounds, t_dev[-1]), flush=True)
        for b in (d_t, d_o) + bufs: b.free()
    bad = []
    legs = {}
    for name, fn, args in (("glob_and_escape", fc.glob_and_escape, (eng, oracle, seed, templates, bad)), ("many_states", fc.many_states, (eng, oracle, seed, ins, templates, bad)),
                           ("host_mirror", fc.host_mirror, (eng, oracle, seed, ins, templates, bad)), ("deep_and_mutate", fc.deep_and_mutate, (eng, oracle, seed, bad))):
        t0 = time.time(); fn(*args); legs[name] = time.time() - t0
        print("  %s %.2f s" % (name, legs[name]), flush=True)
    print("seed %d: %d inserts %d templates | pack %.2f oracle %.2f host %.2f (kernel_ms %.1f, general %d) device r0-r3 %s | %s" % (
        seed, len(ins), n, t_pack, t_or, t_host, got.kernel_ms, got.n_general, " ".join("%.3f" % t for t in t_dev), " ".join("%s %.2f" % kv for kv in legs.items())), flush=True)
