"""Long differential run of the CUDA resolve path against the oracle (a script, not collected by pytest):
    python tests/fuzz_campaign.py [seconds] [first_seed] [--mirror-only]
Every iteration builds one table from a handful of random insert sets and a batch of several thousand templates from the
brace / escape / sentinel-heavy alphabet of tests/casegen.py — short ones, concatenations of several of them with literals in
between (so that tiles see long and dense templates next to empty ones), and plain C4-like ones — and compares bytes, status
and tag of every result through the host-buffer call and through the device-buffer call with 0-3 rescan rounds; the same
strings then serve as keys of random wildcard sweeps (bitmask against the oracle's matcher) and as input of escape / unescape
(against the two-pass replace of interp.rs:149,165), and a slice of the batch is resolved against many perturbed snapshots
of the table in one launch (ie_table_pack_many); finally the JSON-level mirror of the module API (tree walkers, replace_map,
goto_map, wildcard_captures, get_interpdata) runs on random values built from the same strings.
Prints the first mismatches with the inserts that produced them; exits 1 if there were any."""
import os, random, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
from tests import oracle_lib
from tests.casegen import gen_cases, BS
from tests.oracle_lib import KIND_TO_CODE

LIMIT = KIND_TO_CODE["limit"]


def batch(seed):
    rng = random.Random(seed)
    cases = gen_cases(seed * 7919 + 1, rng.choice([300, 1500, 5000]))
    ins = {}
    for c_ins, _ in cases[:rng.randint(1, 8)]:
        ins.update({k: v for k, v in c_ins.items() if not isinstance(v, float)})
    short = [t for _, t in cases]
    templates = list(short)
    for _ in range(len(short) // 3):   # longer / denser templates
        k = rng.randint(2, 12)
        glue = rng.choice(["", " ", " lit ", "." + BS + "}", "x" * rng.randint(1, 90)])
        templates.append(glue.join(rng.choice(short) for _ in range(k)))
    for _ in range(len(short) // 2):   # the plain shape of the hot path
        templates.append("text " * rng.randint(0, 12) + "{k%d}" % rng.randint(1, 3) + " and {a-{n1}} " + "{%s}" % rng.choice(list(ins) or ["a"]))
    templates += ["", "{", "}", BS, "{}", "x" * 5000 + "{a}", "{a}" * 300]
    rng.shuffle(templates)
    return ins, templates


def big_table_batch(seed):
    """A table of hundreds to tens of thousands of keys (probe sequences, long keys, values at every length class
    boundary, numbers as chain links, values that refer to lower-numbered keys) and templates over it."""
    rng = random.Random(seed ^ 0xB16)
    nk = rng.choice([200, 3000, 20000])
    lens = [0, 1, 2, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 255, 256, 257, 1000, 4000]
    ins = {}
    long_keys = []
    for i in range(nk):
        r = rng.random()
        if r < 0.35:
            ins["k%d" % i] = ("v%d " % i) + "abcdefghij"[: rng.randint(0, 10)] * rng.randint(0, 3)
        elif r < 0.50:
            ins["k%d" % i] = "x" * rng.choice(lens)
        elif r < 0.65:
            ins["n%d" % i] = rng.randint(0, nk - 1)                      # chain link: {k{n<i>}}
        elif r < 0.75 and i > 4:
            j, j2 = rng.randint(0, i - 1), rng.randint(0, i - 1)           # rescans, always towards lower numbers
            ins["k%d" % i] = rng.choice(["{k%d}", "see {k%d}.", "{k%d}{k%d}", "a {k{n%d}} b", BS + "{{k%d}" + BS + "}", "{k%d}" + BS])\
                .replace("%d", str(j), 1).replace("%d", str(j2))
        elif r < 0.80:
            lk = "pre-%d-" % i + "y" * rng.choice([20, 60, 100, 250, 300])
            ins[lk] = "long key %d" % i
            long_keys.append(lk)
        elif r < 0.85:
            ins["k%d" % i] = rng.choice([True, False, None, [1, "a", [2]], {"o": i}, [], 0, -i, 10 ** 15 + i])
        elif r < 0.90:
            ins["k%d" % i] = "".join(rng.choice(["{", "}", BS, ".", "a", "〠"]) for _ in range(rng.randint(1, 8)))
        else:
            ins["pre-%d-suf" % i] = "p%d" % i
    keys = list(ins)
    ids = {p: [int(k[len(p):]) for k in keys if k.startswith(p) and k[len(p):].isdigit()] or [0] for p in ("k", "n")}
    # chain links mostly land on keys that exist
    for k in keys:
        if k.startswith("n") and rng.random() < 0.8:
            ins[k] = rng.choice(ids["k"])
            if rng.random() < 0.7:
                ins["pre-%d-suf" % ins[k]] = "P%d" % ins[k]
    templates = []
    for _ in range(rng.choice([2000, 8000])):
        parts = []
        for _ in range(rng.randint(0, 7)):
            r = rng.random()
            miss = rng.random() < 0.03
            i = rng.randint(0, nk - 1) if miss else rng.choice(ids["k"])
            if r < 0.25:   parts.append("{k%d}" % i)
            if not miss: i = rng.choice(ids["n"])
            if r < 0.25:   pass
            elif r < 0.40: parts.append("{k{n%d}}" % i)
            elif r < 0.50: parts.append("{pre-{n%d}-suf}" % i)
            elif r < 0.58: parts.append("{%s}" % rng.choice(keys))
            elif r < 0.62 and long_keys: parts.append("{%s}" % rng.choice(long_keys))
            elif r < 0.63: parts.append("{k{k{n%d}}}" % i)
            elif r < 0.70: parts.append(BS + "{lit" + BS + "}")
            elif r < 0.72: parts.append(rng.choice(["{", "}", "{}", BS, "." + BS + "}"]))
            else:          parts.append("literal text "[: rng.randint(0, 13)] * rng.randint(0, 6))
        t = "".join(parts)
        if rng.random() < 0.08:
            t = "{" + t + "}"
        templates.append(t)
    return ins, templates


def compare(label, seed, ins, templates, status_g, lens_g, get_g, tags_g, out, offs, status, aux, bad):
    for i, t in enumerate(templates):
        if status[i] == LIMIT:
            continue
        w = out[int(offs[i]):int(offs[i + 1])].tobytes()
        if (status_g[i] & 0xFF) == LIMIT and len(w) > 30000:   # the CUDA path caps a result at 64 KiB, the oracle does not
            continue
        ok = (status_g[i] & 0xFF) == (status[i] & 0xFF) and get_g(i) == w
        if ok and status[i] == ie.RES_TYPED and tags_g is not None:
            ok = tags_g[i] == aux[i] >> 28
        if not ok:
            bad.append((label, seed, t, int(status_g[i]), int(status[i]), get_g(i)[:80], w[:80]))
            if len(bad) < 6:
                print("MISMATCH", label, "seed", seed, repr(t)[:200], "gpu", status_g[i] & 0xFF, get_g(i)[:80], "oracle", status[i] & 0xFF, w[:80], flush=True)
                print("  inserts:", repr(ins)[:300], "... (%d keys; rebuild with the seed)" % len(ins), flush=True)


def two_pass(b, mode):
    """interp.rs:149 / :165: two sequential non-overlapping replaces."""
    if mode == 0:
        return b.replace(b"\\{", b"{").replace(b"\\}", b"}")
    return b.replace(b"{", b"\\{").replace(b"}", b"\\}")


def glob_and_escape(eng, oracle, seed, templates, bad):
    """The same batch as keys of wildcard sweeps (1-9 random patterns built from pieces of the keys themselves, with up
    to four '*' runs and pieces around the 32-byte limit of the compiled form) and through escape / unescape."""
    rng = random.Random(seed ^ 0x5EED)
    words = ["persona-%d" % rng.randint(0, 30), "/field-%d" % rng.randint(0, 12), "a", "ab", "-", "/", "x" * 31, "y" * 32, "z" * 33, "é", ""]
    keys = ["".join(rng.choice(words) for _ in range(rng.randint(0, 5))) for _ in range(4000)] + [t for t in templates[:2000] if len(t) < 200]
    ka = ie.Arena.from_strings(keys)
    n_cmp = 0
    for _ in range(6):
        pats = []
        for _ in range(rng.randint(0, 9)):
            pieces = [rng.choice(words + [rng.choice(keys)[:rng.randint(0, 40)]]) for _ in range(rng.randint(1, 5))]
            pats.append(rng.choice(["", "*"]) + rng.choice(["*", "**", "*"]).join(pieces) + rng.choice(["", "*"]))
        pa = ie.Arena.from_strings(pats)
        for invert in (False, True):
            mask, nd = eng.glob_sweep(ka, pa, invert)
            want = oracle.glob_sweep(ka.bytes, ka.offs, pa.bytes, pa.offs, invert, threads=4)
            n_cmp += len(keys)
            if not np.array_equal(mask, want):
                bad.append(("glob", seed, pats, invert))
                if len(bad) < 6:
                    w = int(np.flatnonzero(mask != want)[0])
                    bit = int(mask[w] ^ want[w]); k = w * 32 + (bit & -bit).bit_length() - 1
                    print("MISMATCH glob seed", seed, "invert", invert, "key", repr(keys[k]), "patterns", pats, flush=True)
    raw = [t.encode() for t in templates]
    arena = ie.Arena.from_strings(raw)
    for mode in (0, 1):
        got = eng.escape_batch(arena, mode).strings()
        n_cmp += len(raw)
        for i, g in enumerate(got):
            if g != two_pass(raw[i], mode):
                bad.append(("escape", seed, mode, raw[i]))
                if len(bad) < 6:
                    print("MISMATCH escape mode", mode, "seed", seed, raw[i][:200], g[:200], flush=True)
                break
    return n_cmp


def many_states(eng, oracle, seed, ins, templates, bad):
    """ie_table_pack_many: S perturbed copies of the table (keys dropped, values swapped or replaced), a slice of the batch
    resolved against every one of them in one launch; result s * n + j against the oracle's table for state s.  With few
    templates per state this is the 32-template tile build of the kernel."""
    rng = random.Random(seed ^ 0x57A7E)
    keys = list(ins)
    n_states = rng.choice([2, 5, 33, 200])
    sub = templates[:rng.choice([3, 32, 33, 150, 700])]
    states = []
    for _ in range(n_states):
        st = dict(ins)
        for k in rng.sample(keys, min(len(keys), rng.randint(0, 6))):
            r = rng.random()
            if r < 0.3: del st[k]
            elif r < 0.6: st[k] = ins[rng.choice(keys)]
            else: st[k] = rng.choice(["", "other", "{%s}" % rng.choice(keys), 42, None, ["l", 1]])
        states.append(st)
    if rng.random() < 0.3:
        states[rng.randrange(n_states)] = {}
    packs = [ie.PackedInserts.from_dict(st) for st in states]
    table = eng.pack_many(packs, hhmm="12:34", hhmmss="12:34:56")
    arena = ie.Arena.from_strings(sub)
    got = eng.resolve_batch(table, arena, limits=(4096, 1 << 16))
    n = arena.n
    for si, pk in enumerate(packs):
        out, offs, status, aux = oracle.build_table(pk).resolve_batch(arena.bytes, arena.offs, threads=4, hhmm="12:34", hhmmss="12:34:56")
        compare("states[%d/%d]" % (si, n_states), seed, states[si], sub, got.status_raw[si * n:(si + 1) * n], None,
                lambda j: got.get(si * n + j), got.tags[si * n:(si + 1) * n], out, offs, status, aux, bad)
    return n * n_states


CLOCK = {"hhmm": "12:34", "hhmmss": "12:34:56"}


def host_mirror(eng, oracle, seed, ins, templates, bad, rounds=25):
    """The JSON-level mirror of the reference's module API (ie_host.cpp) against the oracle's: tree walkers over random
    nested values whose keys and strings are templates of the batch, replace_map / goto_map with random wildcard maps."""
    rng = random.Random(seed ^ 0x40057)
    ins = {k: v for k, v in list(ins.items())[:400]}
    short = [t for t in templates if len(t) < 120][:500] or [""]

    def tree(depth):
        r = rng.random()
        if depth == 0 or r < 0.45: return rng.choice(short)
        if r < 0.55: return rng.choice([None, True, 3, -1.5, 10 ** 12])
        if r < 0.75: return [tree(depth - 1) for _ in range(rng.randint(0, 4))]
        d = {rng.choice(short): tree(depth - 1) for _ in range(rng.randint(0, 4))}
        if rng.random() < 0.3:
            d["cmd"] = rng.choice(["print", "goto_map", "replace_map", "serial", "for", "parallel_race", "parallel_wait", "set"])
            if rng.random() < 0.7: d["tasks"] = rng.choice([rng.choice(short), [rng.choice(short), 5], "{%s}" % rng.choice(list(ins) or ["a"])])
        return d

    def check(fn, **kw):
        # self-referential values make the reference's loop explode; the oracle gives up early and the case is skipped
        # (before the engine is asked: it would need minutes to reach its own, much larger limits)
        t0 = time.time()
        want = oracle.call(fn, clock=CLOCK, max_iterations=600, max_bytes=1 << 17, **{k: v for k, v in kw.items() if k != "max_iterations"})
        if want[0] == "err" and want[1]["code"] == LIMIT:
            return 0
        t1 = time.time()
        got = eng.call(fn, clock=CLOCK, **kw)
        if time.time() - t0 > 2.0:
            print("SLOW %s seed %d: oracle %.1f s, engine %.1f s: %s" % (fn, seed, t1 - t0, time.time() - t1, repr({k: v for k, v in kw.items() if k != "inserts"})[:400]), flush=True)
        ok = got == want
        if not ok:
            bad.append((fn, seed, kw))
            if len(bad) < 6:
                print("MISMATCH", fn, "seed", seed, repr({k: v for k, v in kw.items() if k != "inserts"})[:600], "\n  gpu   ", repr(got)[:400], "\n  oracle", repr(want)[:400], flush=True)
        return 1

    n = 0
    for _ in range(rounds):
        v = tree(3)
        n += check("recursive_interpolate", inserts=ins, value=v)
        n += check("recursive_escape", value=v)
        n += check("recursive_unescape", value=v)
        maps = []
        for _ in range(rng.randint(0, 5)):
            k = "".join(rng.choice(["*", "*", rng.choice(short)[:rng.randint(0, 8)], " ", "a", "{%s}" % rng.choice(list(ins) or ["a"])]) for _ in range(rng.randint(0, 4)))
            val = "".join(rng.choice(["{1}", "{2}", "{3}", " ", "z", rng.choice(short)[:12]]) for _ in range(rng.randint(0, 3)))
            maps.append({k: val})
        if rng.random() < 0.3:
            maps.append({"NULL": rng.choice(["nil", "{1}", ""])})
        item = rng.choice(short)
        for rep in (False, True):
            n += check("replace_map", inserts=ins, item=item, wildcard_maps=maps, repeat_until_done=rep)
        n += check("goto_map", inserts=ins, text=item, target_maps=[{k: "T" + v} for m in maps for k, v in m.items()])
        n += check("wildcard_captures", pattern="".join(rng.choice(["*", "a", "b", " ", item[:3]]) for _ in range(rng.randint(0, 6))), text=item)
        n += check("interpolate_inserts", inserts=ins, content=item, max_iterations=4096)
        n += check("get_interpdata", inserts=ins, key=rng.choice(list(ins) + ["", "ARG3", "nope", "HH:MM"]))
    return n


def deep_and_mutate(eng, oracle, seed, bad):
    """(1) Inputs that need more than the default bounds - value chains with text left of every group (one splice frame
    per level), results of hundreds of kilobytes on the general path, thousands of lookups in one template - through the
    host API with default limits: wherever the oracle (given generous bounds) resolves, the engine must too, never
    "limit".  (2) A table patched in place by random set / delete operations against a dict that follows them."""
    rng = random.Random(seed ^ 0xDEE9)
    n = 0
    depth = rng.choice([3, 26, 60, 150, 400, 1500])
    ins = {"A%d" % k: rng.choice(["x%d" % k, "", "lit "]) + "{A%d}" % (k + 1) + rng.choice(["", ".", "!"]) for k in range(1, depth)}
    ins["A%d" % depth] = rng.choice(["end", "", "y" * rng.choice([10, 3000, 70000])])
    ins.update({"big": "ab" * rng.choice([10, 40000, 300000]), "k": "v", "e": ""})
    many = rng.choice([10, 3000, 6000])
    templates = ["v={A1}", "{A1}", "." + BS + "}{A1}", "<{A%d}>" % max(1, depth // 2), "." + BS + "}" + "{k}{e}" * many, "{k}{e}" * many, "." + BS + "}<{big}>{big}",
                 "{big}{big}", "." + BS + "}{A%d}|{big}" % max(1, depth - 2)]
    kind, res = eng.call("interpolate_many", inserts=ins, contents=templates, clock=CLOCK)
    assert kind == "ok"
    for t, r in zip(templates, res):
        want = oracle.call("interpolate_inserts", inserts=ins, content=t, clock=CLOCK, max_iterations=100000, max_bytes=1 << 26)
        got = ("ok", r["ok"]) if "ok" in r else ("err", r["err"])
        n += 1
        if got != want:
            bad.append(("deep", seed, t[:80]))
            if len(bad) < 6:
                print("MISMATCH deep seed", seed, "depth", depth, repr(t)[:120], "\n  gpu   ", repr(got)[:200], "\n  oracle", repr(want)[:200], flush=True)
    # (2) in-place mutation
    cur = {"a": "A", "i": 1, "q-1": "first", "lst": ["x", 2]}
    cur.update({"pad-%d" % k: "p%d" % k for k in range(rng.choice([0, 40, 5000]))})
    table = eng.pack(ie.PackedInserts.from_dict(cur))
    keys = ["a", "b", "c", "i", "q-1", "q-2", "k" * 16, "K" * 17, "a key that is clearly longer than sixteen bytes", "pad-3"]
    values = ["", "v", "x" * 15, "y" * 16, "z" * 17, "long value " * rng.randint(2, 40), 3, -5, ["l", 2], True, None, {"o": 1}]
    tpl = ["{a}", "<{a}|{b}|{c}>", "{%s}" % ("k" * 16), "x{%s}" % ("K" * 17), "{a key that is clearly longer than sixteen bytes}!", "{q-{i}}", "{i}", "{missing}", "{{b}}",
           "{pad-3}{pad-4}", "{lst}"]
    arena = ie.Arena.from_strings(tpl)
    for step in range(25):
        try:
            if rng.random() < 0.7:
                ops = [(rng.choice(keys), rng.choice(values)) for _ in range(rng.randint(1, 5))]
                table.set(ops)
                for k, v in ops:
                    cur[k] = v
            else:
                ks = [rng.choice(keys + ["never-there"]) for _ in range(rng.randint(1, 3))]
                table.delete(ks)
                for k in ks:
                    cur.pop(k, None)
        except ie.EngineError as ex:  # no room left: a fresh pack takes over (the operations that fit were applied)
            assert "pack the snapshot again" in str(ex)
            cur = None
        if cur is None:
            break
        got = eng.resolve_batch(table, arena)
        out, offs, status, aux = oracle.build_table(ie.PackedInserts.from_dict(cur)).resolve_batch(arena.bytes, arena.offs)
        compare("mutate step %d" % step, seed, cur, tpl, got.status_raw, got.lens, got.get, got.tags, out, offs, status, aux, bad)
        n += len(tpl)
    return n


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    for a in sys.argv[1:]:
        if a.startswith("--lib="):  # e.g. --lib=libie_b200_dbg.so: the IE_DEBUG_BOUNDS build of the library
            ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), a[len("--lib="):])
    seconds = float(argv[0]) if argv else 60.0
    seed = first_seed = int(argv[1]) if len(argv) > 1 else 1
    eng, oracle = ie.Engine(0), oracle_lib.load()
    t_end, n_done, bad = time.time() + seconds, 0, []
    leg_s = {}  # seconds per leg (printed at the end: a leg that suddenly dominates is a performance cliff)

    def timed(name, fn, *a, **kw):
        t0 = time.time()
        r = fn(*a, **kw)
        leg_s[name] = leg_s.get(name, 0.0) + time.time() - t0
        return r
    mirror_only = "--mirror-only" in sys.argv   # only the JSON-level leg (cheap per check: for volume there)
    while time.time() < t_end and len(bad) < 20:
        t_batch = time.time()
        ins, templates = batch(seed) if seed % 3 else big_table_batch(seed)
        if mirror_only:
            n_done += host_mirror(eng, oracle, seed, ins, templates, bad, rounds=200)
            seed += 1
            continue
        packed = ie.PackedInserts.from_dict(ins)
        table = eng.pack(packed, hhmm="12:34", hhmmss="12:34:56")
        arena = ie.Arena.from_strings(templates)
        out, offs, status, aux = oracle.build_table(packed).resolve_batch(arena.bytes, arena.offs, threads=8, hhmm="12:34", hhmmss="12:34:56")
        got = eng.resolve_batch(table, arena, limits=(4096, 1 << 16))
        compare("host", seed, ins, templates, got.status_raw, got.lens, got.get, got.tags, out, offs, status, aux, bad)
        # device-buffer call, arenas at odd addresses, every rescan-round setting
        n, nb = arena.n, arena.bytes.nbytes
        cap = int((offs[1:] - offs[:-1]).sum()) * 2 + (1 << 20)
        for rounds in (0, 1, 2, 3):
            shift = (seed + rounds) % 16
            d_t = eng.alloc(nb + 64).upload(np.concatenate([np.zeros(shift, np.uint8), arena.bytes]))
            d_o = eng.alloc((n + 1) * 8).upload(arena.offs)
            bufs = (eng.alloc(cap + 64), eng.alloc(n * 8), eng.alloc(n * 4), eng.alloc(n * 4), eng.alloc(n * 4), eng.alloc(64))
            eng.resolve_batch_device(table, d_t.ptr + shift, d_o.ptr, n, bufs[0].ptr + (shift * 3) % 16, cap, bufs[1].ptr, bufs[2].ptr, bufs[3].ptr,
                                     bufs[4].ptr, bufs[5].ptr, limits=(4096, 1 << 16, 0, 0, rounds))
            eng.sync()
            o = bufs[0].download(np.uint8, cap + 64)[(shift * 3) % 16:]
            oo, ol, st = bufs[1].download(np.uint64, n), bufs[2].download(np.uint32, n), bufs[3].download(np.int32, n)
            compare("device r%d" % rounds, seed, ins, templates, st, ol, lambda i: o[int(oo[i]):int(oo[i]) + int(ol[i])].tobytes(), None, out, offs, status,
                    aux, bad)
            for b in (d_t, d_o) + bufs:
                b.free()
        leg_s["resolve"] = leg_s.get("resolve", 0.0) + time.time() - t_batch
        n_done += len(templates) * 5 + timed("glob_and_escape", glob_and_escape, eng, oracle, seed, templates, bad)
        n_done += timed("many_states", many_states, eng, oracle, seed, ins, templates, bad)
        n_done += timed("host_mirror", host_mirror, eng, oracle, seed, ins, templates, bad)
        n_done += timed("deep_and_mutate", deep_and_mutate, eng, oracle, seed, bad)
        if "--progress" in sys.argv:
            print("seed %d done at %.1f s: %s" % (seed, time.time() - (t_end - seconds), ", ".join("%s %.1f" % kv for kv in sorted(leg_s.items()))), flush=True)
        seed += 1
    print("fuzz campaign: %d results compared over %d batches, %d mismatches" % (n_done, seed - first_seed, len(bad)))
    print("seconds per leg: " + ", ".join("%s %.1f" % kv for kv in sorted(leg_s.items())))
    for name in ("ie_debug_bound_violations", "ie_debug_bound_violations_small", "ie_debug_bound_violations_fused"):
        if hasattr(eng.lib, name):  # IE_DEBUG_BOUNDS build: every tile-table index was checked against its capacity
            import ctypes
            v = (ctypes.c_ulonglong * 4)()
            getattr(eng.lib, name)(v)
            print("%s: %d violations (first: line %d, index %d, capacity %d)" % (name, v[0], v[1], v[2], v[3]))
            if v[0]:
                bad.append(("bounds", name))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
