"""interpolation_engine_b200 — B200-native batched `{key}` interpolation (ctypes binding of the C ABI).

The package is a thin host layer over ``libie_b200.so`` (include/ie_b200.h): CUDA kernels for sm_100a
that resolve many independent templates against one packed inserts snapshot, plus the escape /
unescape and wildcard-delete kernels.  There is no CPU fallback: importing works anywhere (so the
ABI can be inspected), but creating an :class:`Engine` without the built library or without a CUDA
device raises.

Reference being replaced: rust-project/src/interp.rs and rust-project/src/runtime.rs:1198-1239,
1633-1647 of tillfalko/interpolation-engine.
"""
import ctypes
import time
import json
import os

import numpy as np

__all__ = ["Engine", "Table", "PackedInserts", "Arena", "EngineError", "load_library", "LIB_PATH", "resolve_batch_multi",
           "RES_STRING", "RES_TYPED", "RES_UNEVEN", "RES_UNSUPPORTED", "RES_EMPTY_KEY", "RES_ARG_MISSING",
           "RES_NOT_FOUND", "RES_PANIC", "RES_LIMIT", "STATUS_NAMES"]

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libie_b200.so")

RES_STRING, RES_TYPED, RES_UNEVEN, RES_UNSUPPORTED, RES_EMPTY_KEY, RES_ARG_MISSING, RES_NOT_FOUND, RES_PANIC, RES_LIMIT = range(9)
STATUS_NAMES = {0: "string", 1: "typed", 2: "uneven", 3: "unsupported", 4: "empty", 5: "arg", 6: "not_found",
                7: "panic", 8: "limit", 9: "io"}
TAG_NULL, TAG_BOOL, TAG_NUMBER, TAG_STRING, TAG_ARRAY, TAG_OBJECT = range(6)
AUX_NONE = 0xFFFFFFFF  # IE_AUX_NONE


class EngineError(RuntimeError):
    pass


class _BatchInfo(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint64), ("out_bytes", ctypes.c_uint64), ("n_general", ctypes.c_uint64),
                ("n_limit", ctypes.c_uint64), ("kernel_ms", ctypes.c_float)]


class _Result(ctypes.Structure):
    _fields_ = [("out", ctypes.c_void_p), ("out_offs", ctypes.c_void_p), ("out_lens", ctypes.c_void_p),
                ("status", ctypes.c_void_p), ("aux", ctypes.c_void_p), ("info", _BatchInfo)]


ALL_STATES = 0xFFFFFFFF  # IE_ALL_STATES


class _ShardResult(ctypes.Structure):
    _fields_ = [("first", ctypes.c_uint64), ("n", ctypes.c_uint64), ("res", _Result), ("status", ctypes.c_int), ("error", ctypes.c_char * 160)]


class _Limits(ctypes.Structure):
    _fields_ = [("max_expansions", ctypes.c_uint32), ("max_result_bytes", ctypes.c_uint32), ("avg_template_bytes", ctypes.c_uint32),
                ("avg_template_groups", ctypes.c_uint32), ("rescan_rounds", ctypes.c_uint32)]


# every symbol include/ie_b200.h declares: name -> (restype, argtypes)
_vp, _u64, _u32, _i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
ABI = {
    "ie_last_error": (ctypes.c_char_p, []),
    "ie_device_count": (_i, []),
    "ie_engine_create": (_i, [_i, ctypes.POINTER(_vp)]),
    "ie_engine_destroy": (None, [_vp]),
    "ie_engine_stream": (_vp, [_vp]),
    "ie_engine_sync": (_i, [_vp]),
    "ie_table_pack": (_i, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "ie_table_pack_many": (_i, [_vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "ie_table_states": (_u32, [_vp]),
    "ie_table_build_ms": (ctypes.c_double, [_vp]),
    "ie_table_set": (_i, [_vp, _vp, _u32, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ie_table_delete": (_i, [_vp, _vp, _u32, _u64, _vp, _vp]),
    "ie_table_free": (None, [_vp]),
    "ie_table_device_bytes": (_u64, [_vp]),
    "ie_resolve_batch": (_i, [_vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_Limits), ctypes.POINTER(_Result)]),
    "ie_resolve_batch_multi": (_i, [_vp, _vp, _u32, _vp, _vp, _u64, ctypes.POINTER(_Limits), _vp]),
    "ie_shards_gather": (_i, [_vp, _u32, _vp, _u64, _vp, _vp, _vp, ctypes.POINTER(_u64)]),
    "ie_resolve_batch_device": (_i, [_vp, _vp, _vp, _vp, _u64, ctypes.POINTER(_Limits), _vp, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "ie_lookup_batch": (_i, [_vp, _vp, _vp, _vp, _u64, _vp, _vp]),
    "ie_escape_batch": (_i, [_vp, _i, _vp, _vp, _u64, ctypes.POINTER(_vp), ctypes.POINTER(_vp)]),
    "ie_escape_batch_device": (_i, [_vp, _i, _vp, _vp, _u64, _u64, _vp, _u64, _vp, _vp]),
    "ie_glob_sweep": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _u32, _i, _vp, ctypes.POINTER(_u64)]),
    "ie_glob_sweep_device": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _u32, _i, _vp, _vp, _vp]),
    "ie_glob_first_match": (_i, [_vp, _vp, _vp, _u64, _vp, _vp, _u32, _vp]),
    "ie_device_alloc": (_i, [_vp, _u64, ctypes.POINTER(_vp)]),
    "ie_device_free": (None, [_vp, _vp]),
    "ie_copy_to_device": (_i, [_vp, _vp, _vp, _u64]),
    "ie_copy_to_host": (_i, [_vp, _vp, _vp, _u64]),
    "ie_host_alloc": (_i, [_u64, ctypes.POINTER(_vp)]),
    "ie_host_alloc_wc": (_i, [_u64, ctypes.POINTER(_vp)]),
    "ie_host_free": (None, [_vp]),
    "ie_call_json": (_i, [_vp, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "ie_free": (None, [_vp]),
}

_lib = None


def load_library():
    """Loads libie_b200.so (built in-tree by __graft_entry__.build()); fails loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EngineError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)  # AttributeError = the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


class Arena:
    """A string arena: `bytes` (uint8) + `offs` (uint64, n+1 entries)."""

    def __init__(self, data, offs):
        self.bytes = np.ascontiguousarray(data, dtype=np.uint8)
        self.offs = np.ascontiguousarray(offs, dtype=np.uint64)

    @classmethod
    def from_strings(cls, strings):
        enc = [s.encode("utf-8") if isinstance(s, str) else bytes(s) for s in strings]
        offs = np.zeros(len(enc) + 1, dtype=np.uint64)
        if enc:
            offs[1:] = np.cumsum([len(b) for b in enc], dtype=np.uint64)
        return cls(np.frombuffer(b"".join(enc), dtype=np.uint8), offs)

    @property
    def n(self):
        return len(self.offs) - 1

    def get(self, i):
        return self.bytes[int(self.offs[i]):int(self.offs[i + 1])].tobytes()

    def strings(self):
        return [self.get(i) for i in range(self.n)]


def render_value(v):
    """value_to_string (interp.rs:314-322) for Python-side packing of test / bench inserts.
    Floats are deliberately unsupported here (serde_json's f64 layout lives in the C++ host layer:
    use Engine.call for inserts that hold floats)."""
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, list):
        return "".join(render_value(x) for x in v)
    if isinstance(v, float):
        raise TypeError("render_value: floats must go through Engine.call (C++ renderer)")
    return json.dumps(v, separators=(",", ":"), sort_keys=True, ensure_ascii=False)


def tag_of(v):
    if v is None:
        return TAG_NULL
    if isinstance(v, bool):
        return TAG_BOOL
    if isinstance(v, (int, float)):
        return TAG_NUMBER
    if isinstance(v, str):
        return TAG_STRING
    if isinstance(v, list):
        return TAG_ARRAY
    return TAG_OBJECT


class PackedInserts:
    """An inserts snapshot in the packed arena form ie_table_pack takes."""

    def __init__(self, keys, key_offs, vals, val_offs, tags):
        self.keys = np.ascontiguousarray(keys, dtype=np.uint8)
        self.key_offs = np.ascontiguousarray(key_offs, dtype=np.uint64)
        self.vals = np.ascontiguousarray(vals, dtype=np.uint8)
        self.val_offs = np.ascontiguousarray(val_offs, dtype=np.uint64)
        self.tags = np.ascontiguousarray(tags, dtype=np.uint8)
        self.n = len(self.tags)

    @classmethod
    def from_dict(cls, inserts):
        items = sorted(inserts.items(), key=lambda kv: kv[0].encode("utf-8"))  # BTreeMap order
        k = Arena.from_strings([key for key, _ in items])
        v = Arena.from_strings([render_value(val) for _, val in items])
        return cls(k.bytes, k.offs, v.bytes, v.offs, np.array([tag_of(val) for _, val in items], dtype=np.uint8))


class Table:
    def __init__(self, engine, handle, packed):
        self.engine, self.handle, self.packed = engine, handle, packed

    @property
    def device_bytes(self):
        return int(self.engine.lib.ie_table_device_bytes(self.handle))

    @property
    def build_ms(self):
        """Device time of the build kernels (0.0 for a host-built table)."""
        return float(self.engine.lib.ie_table_build_ms(self.handle))

    def set(self, items, state=None, entries=None):
        """set_interpdata (interp.rs:139) in place, for each (key, value) of `items` (dict or list of pairs), in order.
        `state`: snapshot of a pack_many table (None = every snapshot)."""
        pairs = list(items.items()) if isinstance(items, dict) else list(items)
        k = Arena.from_strings([key for key, _ in pairs])
        v = Arena.from_strings([render_value(val) for _, val in pairs])
        tags = np.array([tag_of(val) for _, val in pairs], dtype=np.uint8)
        ent = np.ascontiguousarray(entries, dtype=np.uint32) if entries is not None else None
        self.engine._check(self.engine.lib.ie_table_set(self.engine.handle, self.handle, ALL_STATES if state is None else int(state), len(pairs),
                                                        _ptr(k.bytes), _ptr(k.offs), _ptr(v.bytes), _ptr(v.offs), _ptr(tags),
                                                        _ptr(ent) if ent is not None else None))

    def delete(self, keys, state=None):
        """delete_interpdata (interp.rs:143) in place."""
        k = Arena.from_strings(list(keys))
        self.engine._check(self.engine.lib.ie_table_delete(self.engine.handle, self.handle, ALL_STATES if state is None else int(state), k.n,
                                                           _ptr(k.bytes), _ptr(k.offs)))

    def free(self):
        if self.handle:
            self.engine.lib.ie_table_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BatchResult:
    """Host copy of one resolved batch."""

    def __init__(self, out, offs, lens, status, aux, info):
        self.out, self.offs, self.lens, self.status_raw, self.aux = out, offs, lens, status, aux
        self.status = status & 0xFF
        self.tags = (status >> 8) & 0xFF
        self.out_bytes, self.n_general, self.kernel_ms = info

    def get(self, i):
        o = int(self.offs[i])
        return self.out[o:o + int(self.lens[i])].tobytes()


def resolve_batch_multi(engines, tables, templates, limits=None, gather=True):
    """ie_resolve_batch_multi: ONE host batch cut into contiguous shards, shard g through engines[g] / tables[g] (one
    host thread per engine inside the C call).  Returns (shards, gathered): the per-shard BatchResults and, with
    gather=True, the host gather of ie_shards_gather as (out, offs[n+1], status, aux)."""
    arena = templates if isinstance(templates, Arena) else Arena.from_strings(templates)
    lib = engines[0].lib
    G = len(engines)
    eh = (ctypes.c_void_p * G)(*[e.handle for e in engines])
    th = (ctypes.c_void_p * G)(*[t.handle for t in tables])
    shards = (_ShardResult * G)()
    lim = _Limits(*limits) if limits else None
    engines[0]._check(lib.ie_resolve_batch_multi(eh, th, G, _ptr(arena.bytes), _ptr(arena.offs), arena.n, ctypes.byref(lim) if lim else None, shards))

    def view(p, dtype, count):
        if not count:
            return np.zeros(0, dtype=dtype)
        return np.frombuffer((ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(p), dtype=dtype, count=count).copy()
    per = []
    for sh in shards:
        n, ob = int(sh.n), int(sh.res.info.out_bytes)
        per.append((int(sh.first), BatchResult(view(sh.res.out, np.uint8, ob), view(sh.res.out_offs, np.uint64, n), view(sh.res.out_lens, np.uint32, n),
                                               view(sh.res.status, np.int32, n), view(sh.res.aux, np.uint32, n),
                                               (ob, int(sh.res.info.n_general), float(sh.res.info.kernel_ms)))))
    gathered = None
    if gather:
        total = ctypes.c_uint64(0)
        offs = np.zeros(arena.n + 1, dtype=np.uint64)
        status = np.zeros(max(arena.n, 1), dtype=np.int32)
        aux = np.zeros(max(arena.n, 1), dtype=np.uint32)
        cap = sum(int(b.lens.sum()) for _, b in per)
        out = np.zeros(max(cap, 1), dtype=np.uint8)
        engines[0]._check(lib.ie_shards_gather(shards, G, _ptr(out), cap, _ptr(offs), _ptr(status), _ptr(aux), ctypes.byref(total)))
        gathered = (out[:int(total.value)], offs, status[:arena.n], aux[:arena.n])
    return per, gathered


class DeviceBuffer:
    def __init__(self, engine, nbytes):
        self.engine, self.nbytes = engine, int(nbytes)
        p = ctypes.c_void_p()
        engine._check(engine.lib.ie_device_alloc(engine.handle, self.nbytes, ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.engine._check(self.engine.lib.ie_copy_to_device(self.engine.handle, self.ptr, _ptr(arr), arr.nbytes))
        return self

    def download(self, dtype, count):
        out = np.empty(count, dtype=dtype)
        assert out.nbytes <= self.nbytes
        self.engine._check(self.engine.lib.ie_copy_to_host(self.engine.handle, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            self.engine.lib.ie_device_free(self.engine.handle, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One engine per (process, GPU).  Raises EngineError when no CUDA device is usable."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        st = self.lib.ie_engine_create(int(device), ctypes.byref(h))
        if st != 0:
            raise EngineError(self.lib.ie_last_error().decode())
        self.handle = h
        self.device = int(device)

    def _check(self, st):
        if st != 0:
            raise EngineError(f"ie status {st}: {self.lib.ie_last_error().decode()}")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ie_engine_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self):
        return self.lib.ie_engine_stream(self.handle)

    def sync(self):
        self._check(self.lib.ie_engine_sync(self.handle))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    # ---- inserts ---------------------------------------------------------------------------
    def pack(self, inserts, hhmm=None, hhmmss=None):
        packed = inserts if isinstance(inserts, PackedInserts) else PackedInserts.from_dict(inserts)
        h = ctypes.c_void_p()
        self._check(self.lib.ie_table_pack(self.handle, packed.n, _ptr(packed.keys), _ptr(packed.key_offs), _ptr(packed.vals),
                                           _ptr(packed.val_offs), _ptr(packed.tags),
                                           hhmm.encode() if hhmm else None, hhmmss.encode() if hhmmss else None, ctypes.byref(h)))
        return Table(self, h, packed)

    def pack_many(self, states, hhmm=None, hhmmss=None):
        """Packs a list of snapshots (dicts or PackedInserts) into ONE table; resolve_batch on it resolves every
        template against every snapshot (result index = snapshot * n_templates + template)."""
        packs = [st if isinstance(st, PackedInserts) else PackedInserts.from_dict(st) for st in states]
        state_offs = np.zeros(len(packs) + 1, dtype=np.uint64)
        state_offs[1:] = np.cumsum([pk.n for pk in packs], dtype=np.uint64)
        kb = np.concatenate([pk.keys for pk in packs]) if packs else np.zeros(0, np.uint8)
        vb = np.concatenate([pk.vals for pk in packs]) if packs else np.zeros(0, np.uint8)
        tags = np.concatenate([pk.tags for pk in packs]) if packs else np.zeros(0, np.uint8)

        def cat_offs(arrs):
            out = np.zeros(int(state_offs[-1]) + 1, dtype=np.uint64)
            base, pos = 0, 0
            for a in arrs:
                m = len(a) - 1
                out[pos:pos + m + 1] = a + np.uint64(base)
                base += int(a[-1])
                pos += m
            return out
        ko, vo = cat_offs([pk.key_offs for pk in packs]), cat_offs([pk.val_offs for pk in packs])
        merged = PackedInserts(kb, ko, vb, vo, tags)
        h = ctypes.c_void_p()
        t0 = time.perf_counter()
        self._check(self.lib.ie_table_pack_many(self.handle, len(packs), _ptr(state_offs), _ptr(kb), _ptr(ko), _ptr(vb), _ptr(vo), _ptr(tags),
                                                hhmm.encode() if hhmm else None, hhmmss.encode() if hhmmss else None, ctypes.byref(h)))
        t = Table(self, h, merged)
        t.n_states = len(packs)
        t.pack_call_s = time.perf_counter() - t0  # the C call alone: image build on host threads + one upload
        return t

    # ---- interpolate_inserts, batched (interp.rs:31-89) ----------------------------------------
    def resolve_batch(self, table, templates, limits=None):
        arena = templates if isinstance(templates, Arena) else Arena.from_strings(templates)
        res = _Result()
        lim = _Limits(*limits) if limits else None
        self._check(self.lib.ie_resolve_batch(self.handle, table.handle, _ptr(arena.bytes), _ptr(arena.offs), arena.n,
                                              ctypes.byref(lim) if lim else None, ctypes.byref(res)))
        n, ob = arena.n * getattr(table, "n_states", 1), int(res.info.out_bytes)

        def view(p, dtype, count):
            if not count:
                return np.zeros(0, dtype=dtype)
            buf = (ctypes.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()
        return BatchResult(view(res.out, np.uint8, ob), view(res.out_offs, np.uint64, n), view(res.out_lens, np.uint32, n),
                           view(res.status, np.int32, n), view(res.aux, np.uint32, n),
                           (ob, int(res.info.n_general), float(res.info.kernel_ms)))

    def resolve_batch_device(self, table, d_tmpl, d_offs, n, d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, d_info,
                             stream=None, limits=None):
        lim = _Limits(*limits) if limits else None
        self._check(self.lib.ie_resolve_batch_device(self.handle, table.handle, d_tmpl, d_offs, n, ctypes.byref(lim) if lim else None,
                                                     d_out, out_cap, d_out_offs, d_out_lens, d_status, d_aux, d_info, stream))

    # ---- escape / unescape (interp.rs:147-177) -----------------------------------------------
    def escape_batch(self, strings, mode):
        arena = strings if isinstance(strings, Arena) else Arena.from_strings(strings)
        ob, oo = ctypes.c_void_p(), ctypes.c_void_p()
        self._check(self.lib.ie_escape_batch(self.handle, int(mode), _ptr(arena.bytes), _ptr(arena.offs), arena.n,
                                             ctypes.byref(ob), ctypes.byref(oo)))
        offs = np.frombuffer((ctypes.c_char * ((arena.n + 1) * 8)).from_address(oo.value), dtype=np.uint64).copy()
        total = int(offs[-1])
        data = np.frombuffer((ctypes.c_char * total).from_address(ob.value), dtype=np.uint8).copy() if total else np.zeros(0, np.uint8)
        return Arena(data, offs)

    # ---- wildcard delete sweep (runtime.rs:1198-1239, 1633-1647) --------------------------------
    def glob_sweep(self, keys, patterns, invert=False):
        ka = keys if isinstance(keys, Arena) else Arena.from_strings(keys)
        pa = patterns if isinstance(patterns, Arena) else Arena.from_strings(patterns)
        mask = np.zeros((ka.n + 31) // 32 + 1, dtype=np.uint32)
        nd = ctypes.c_uint64(0)
        self._check(self.lib.ie_glob_sweep(self.handle, _ptr(ka.bytes), _ptr(ka.offs), ka.n, _ptr(pa.bytes), _ptr(pa.offs), pa.n,
                                           int(bool(invert)), _ptr(mask), ctypes.byref(nd)))
        return mask[:(ka.n + 31) // 32], int(nd.value)

    def glob_first_match(self, keys, patterns):
        """Index of the first pattern matching each key (-1 = none): goto_map / replace_map selection."""
        ka = keys if isinstance(keys, Arena) else Arena.from_strings(keys)
        pa = patterns if isinstance(patterns, Arena) else Arena.from_strings(patterns)
        first = np.zeros(max(ka.n, 1), dtype=np.uint32)
        self._check(self.lib.ie_glob_first_match(self.handle, _ptr(ka.bytes), _ptr(ka.offs), ka.n, _ptr(pa.bytes), _ptr(pa.offs), pa.n, _ptr(first)))
        return first[:ka.n].view(np.int32).copy() if ka.n else np.zeros(0, np.int32)

    def lookup_batch(self, table, keys):
        """get_interpdata's map probe for a batch of literal keys (interp.rs:118-120): (tags, entries); tag -1 and
        entry AUX_NONE on a miss, entries n / n + 1 = the clock keys."""
        ka = keys if isinstance(keys, Arena) else Arena.from_strings(keys)
        tags = np.full(max(ka.n, 1), -1, dtype=np.int32)
        entries = np.zeros(max(ka.n, 1), dtype=np.uint32)
        self._check(self.lib.ie_lookup_batch(self.handle, table.handle, _ptr(ka.bytes), _ptr(ka.offs), ka.n, _ptr(tags), _ptr(entries)))
        return tags[:ka.n], entries[:ka.n]

    def glob_sweep_device(self, d_keys, d_key_offs, n, patterns, invert, d_mask, d_n_deleted, stream=None):
        pa = patterns if isinstance(patterns, Arena) else Arena.from_strings(patterns)
        self._check(self.lib.ie_glob_sweep_device(self.handle, d_keys, d_key_offs, n, _ptr(pa.bytes), _ptr(pa.offs), pa.n,
                                                  int(bool(invert)), d_mask, d_n_deleted, stream))

    # ---- JSON-level mirror of interp.rs ---------------------------------------------------------
    def call(self, fn, **kw):
        """Returns ("ok", value) or ("err", {"code", "message", "payload"})."""
        kw["fn"] = fn
        blob = json.dumps(kw).encode("utf-8")
        out, n = ctypes.c_void_p(), ctypes.c_size_t(0)
        self._check(self.lib.ie_call_json(self.handle, blob, len(blob), ctypes.byref(out), ctypes.byref(n)))
        try:
            res = json.loads(ctypes.string_at(out.value, n.value).decode("utf-8"))
        finally:
            self.lib.ie_free(out)
        return ("ok", res["ok"]) if "ok" in res else ("err", res["err"])
