"""ctypes loader for the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module; the product package never does.
"""
import ctypes
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "liboracle.so")

STATUS_NAMES = {0: "string", 1: "typed", 2: "uneven", 3: "unsupported", 4: "empty", 5: "arg",
                6: "not_found", 7: "panic", 8: "limit", 9: "io"}
KIND_TO_CODE = {v: k for k, v in STATUS_NAMES.items()}


def build():
    srcs = [os.path.join(ROOT, "oracle", f) for f in ("oracle_capi.cpp", "oracle_interp.hpp", "oracle_value.hpp")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_call_json.restype = ctypes.c_void_p
        lib.orc_call_json.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        lib.orc_free.argtypes = [ctypes.c_void_p]
        lib.orc_table_build.restype = ctypes.c_void_p
        lib.orc_table_build.argtypes = [ctypes.c_uint64] + [ctypes.c_void_p] * 5
        lib.orc_table_free.argtypes = [ctypes.c_void_p]
        lib.orc_resolve_batch.restype = ctypes.c_int
        lib.orc_resolve_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int,
                                          ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p),
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.orc_glob_sweep.restype = ctypes.c_int
        lib.orc_glob_sweep.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

        lib.orc_compare_ragged.restype = ctypes.c_uint64
        lib.orc_compare_ragged.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_uint64]

    def first_mismatch(self, a, a_offs, b, b_offs, lens):
        """Index of the first string that differs between two (arena, offsets) pairs sharing `lens`, or None."""
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        a_offs = np.ascontiguousarray(a_offs, dtype=np.uint64)
        b_offs = np.ascontiguousarray(b_offs, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        i = self.lib.orc_compare_ragged(a.ctypes.data, a_offs.ctypes.data, b.ctypes.data, b_offs.ctypes.data, lens.ctypes.data, len(lens))
        return None if i == len(lens) else int(i)

    def call(self, fn, **kw):
        """Returns ("ok", value) or ("err", {"code", "message", "payload"})."""
        kw["fn"] = fn
        blob = json.dumps(kw).encode("utf-8")
        n = ctypes.c_size_t(0)
        p = self.lib.orc_call_json(blob, len(blob), ctypes.byref(n))
        try:
            res = json.loads(ctypes.string_at(p, n.value).decode("utf-8"))
        finally:
            self.lib.orc_free(p)
        if "ok" in res:
            return "ok", res["ok"]
        return "err", res["err"]

    def build_table(self, packed):
        """packed: PackedInserts-like (keys, key_offs, vals, val_offs, tags as numpy arrays)."""
        return OracleTable(self, packed)

    def glob_sweep(self, keys, key_offs, pats, pat_offs, invert, threads=1):
        n = len(key_offs) - 1
        mask = np.zeros((n + 31) // 32, dtype=np.uint32)
        keys = np.ascontiguousarray(keys)
        pats = np.ascontiguousarray(pats)
        self.lib.orc_glob_sweep(keys.ctypes.data, key_offs.ctypes.data, n, pats.ctypes.data, pat_offs.ctypes.data,
                                len(pat_offs) - 1, int(invert), threads, mask.ctypes.data)
        return mask


class OracleTable:
    def __init__(self, orc, packed):
        self.orc = orc
        self.packed = packed  # keep arrays alive
        self.h = orc.lib.orc_table_build(packed.n, packed.keys.ctypes.data, packed.key_offs.ctypes.data,
                                         packed.vals.ctypes.data, packed.val_offs.ctypes.data, packed.tags.ctypes.data)

    def resolve_batch(self, tmpl, offs, threads=1, hhmm=None, hhmmss=None):
        n = len(offs) - 1
        tmpl = np.ascontiguousarray(tmpl)
        out_offs = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        aux = np.zeros(n, dtype=np.uint32)
        arena = ctypes.c_void_p()
        self.orc.lib.orc_resolve_batch(self.h, tmpl.ctypes.data, offs.ctypes.data, n, threads,
                                       hhmm.encode() if hhmm else None, hhmmss.encode() if hhmmss else None,
                                       ctypes.byref(arena), out_offs.ctypes.data, status.ctypes.data, aux.ctypes.data)
        total = int(out_offs[n])
        out = np.frombuffer(ctypes.string_at(arena.value, total), dtype=np.uint8).copy() if total else np.zeros(0, np.uint8)
        self.orc.lib.orc_free(arena)
        return out, out_offs, status, aux

    def __del__(self):
        try:
            self.orc.lib.orc_table_free(self.h)
        except Exception:
            pass


def load():
    build()
    return Oracle(ctypes.CDLL(LIB))
