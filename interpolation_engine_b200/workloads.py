"""Synthetic workloads of BASELINE.json / SURVEY.md §8(d), generated with numpy (seeded, vectorised).

C4: 1 Mi templates x 65 536-insert state, depth-3 nesting  -> c4_state(), c4_templates()
C5: 10 M keys `persona-<p>/field-<f>` and 64 wildcard sets   -> c5_keys(), c5_pattern_sets()
C3: text_adventure-derived templates over cloned states     -> c3_state(), c3_templates()
"""
import json
import os

import numpy as np

from . import Arena, PackedInserts, TAG_NUMBER, TAG_STRING

ALPHABET = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ,.", dtype=np.uint8)
_DIGITS = None


def _digit_table(limit=100000):
    """[limit, 5] ASCII digits (left aligned) and the digit count of every integer < limit."""
    global _DIGITS
    if _DIGITS is None or _DIGITS[0].shape[0] < limit:
        v = np.arange(limit, dtype=np.int64)
        nd = np.where(v < 10, 1, np.where(v < 100, 2, np.where(v < 1000, 3, np.where(v < 10000, 4, np.where(v < 100000, 5, 6))))).astype(np.int64)
        tab = np.zeros((limit, 6), dtype=np.uint8)
        for k in range(6):
            # k-th digit from the left = (v // 10^(nd-1-k)) % 10 where k < nd
            p = np.maximum(nd - 1 - k, 0)
            tab[:, k] = (v // (10 ** p)) % 10 + 48
        _DIGITS = (tab, nd)
    return _DIGITS


class _Builder:
    """Assembles n strings from per-string pieces without Python-level loops over strings."""

    def __init__(self, n, rng):
        self.n, self.rng = n, rng
        self.pieces = []  # (kind, lengths, payload)

    def literal(self, lengths):
        self.pieces.append(("rand", np.asarray(lengths, dtype=np.int64), None))

    def const(self, text, mask=None):
        b = np.frombuffer(text.encode() if isinstance(text, str) else text, dtype=np.uint8)
        lens = np.full(self.n, len(b), dtype=np.int64)
        if mask is not None:
            lens = lens * mask.astype(np.int64)
        self.pieces.append(("const", lens, b))

    def number(self, values, mask=None):
        tab, nd = _digit_table(max(100000, int(values.max()) + 1) if len(values) else 100000)
        lens = nd[values]
        if mask is not None:
            lens = lens * mask.astype(np.int64)
        self.pieces.append(("num", lens, np.asarray(values, dtype=np.int64)))

    def build(self):
        lens = np.stack([p[1] for p in self.pieces], axis=1)  # [n, pieces]
        total_per = lens.sum(axis=1)
        offs = np.zeros(self.n + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(total_per, dtype=np.uint64)
        total = int(offs[-1])
        arena = ALPHABET[self.rng.integers(0, len(ALPHABET), size=total, dtype=np.uint8)] if total else np.zeros(0, np.uint8)
        starts = offs[:-1].astype(np.int64)[:, None] + np.concatenate([np.zeros((self.n, 1), np.int64), np.cumsum(lens, axis=1)[:, :-1]], axis=1)
        tab, _ = _digit_table()
        for j, (kind, plen, payload) in enumerate(self.pieces):
            st = starts[:, j]
            if kind == "const":
                sel = plen > 0
                for k in range(len(payload)):
                    arena[st[sel] + k] = payload[k]
            elif kind == "num":
                for k in range(6):
                    sel = plen > k
                    if not sel.any():
                        break
                    arena[st[sel] + k] = tab[payload[sel], k]
        return Arena(arena, offs)


def c4_state(seed=0xC4, n_slot=16384, n_idx=16384, n_q=32768):
    """65 536 inserts: slot-<a> -> Number, idx-<b> -> Number, q-<c> -> string U[16,112] (2 % hold `\\{x\\}`)."""
    rng = np.random.default_rng(seed)
    n = n_slot + n_idx + n_q
    kb = _Builder(n, rng)
    kind = np.concatenate([np.zeros(n_slot, np.int64), np.ones(n_idx, np.int64), np.full(n_q, 2, np.int64)])
    ident = np.concatenate([np.arange(n_slot), np.arange(n_idx), np.arange(n_q)]).astype(np.int64)
    kb.const("slot-", kind == 0)
    kb.const("idx-", kind == 1)
    kb.const("q-", kind == 2)
    kb.number(ident)
    keys = kb.build()
    vb = _Builder(n, rng)
    slot_vals = rng.integers(0, n_idx, size=n_slot)
    idx_vals = rng.integers(0, n_q, size=n_idx)
    num = np.concatenate([slot_vals, idx_vals, np.zeros(n_q, np.int64)]).astype(np.int64)
    vb.number(num, kind < 2)
    qlen = np.concatenate([np.zeros(n_slot + n_idx, np.int64), rng.integers(16, 113, size=n_q)])
    esc = np.concatenate([np.zeros(n_slot + n_idx, bool), rng.random(n_q) < 0.02])
    head = (qlen * rng.random(n)).astype(np.int64)
    head = np.where(esc, np.minimum(head, np.maximum(qlen - 5, 0)), qlen)
    vb.literal(head)
    vb.const("\\{x\\}", esc)
    vb.literal(np.where(esc, np.maximum(qlen - 5 - head, 0), 0))
    vals = vb.build()
    tags = np.where(kind < 2, TAG_NUMBER, TAG_STRING).astype(np.uint8)
    return PackedInserts(keys.bytes, keys.offs, vals.bytes, vals.offs, tags)


def c4_templates(n=1 << 20, seed=0xC4, n_slot=16384, n_q=32768, start=0):
    """lit(U[16,96]) + "{q-{idx-{slot-A}}}" + lit(U[16,96]) + "{q-C}" + lit(U[0,64]); 1 % carry a
    literal `\\{not_a_key\\}`; 0.1 % end with `{missing-<n>}`.  `start` offsets the stream so that
    rank r of a sharded run generates templates [start, start + n) of the same global sequence."""
    rng = np.random.default_rng([seed, start])
    b = _Builder(n, rng)
    b.literal(rng.integers(16, 97, size=n))
    b.const("{q-{idx-{slot-")
    b.number(rng.integers(0, n_slot, size=n))
    b.const("}}}")
    b.literal(rng.integers(16, 97, size=n))
    b.const("{q-")
    b.number(rng.integers(0, n_q, size=n))
    b.const("}")
    b.literal(rng.integers(0, 65, size=n))
    b.const("\\{not_a_key\\}", rng.random(n) < 0.01)
    miss = rng.random(n) < 0.001
    b.const("{missing-", miss)
    b.number(rng.integers(0, 100000, size=n), miss)
    b.const("}", miss)
    return b.build()


def c5_keys(n_persona=100000, n_field=100):
    """`persona-<p>/field-<f>` for all (p, f), in (p, f) order (n_persona * n_field keys)."""
    n = n_persona * n_field
    rng = np.random.default_rng(0xC5)
    b = _Builder(n, rng)
    p = np.repeat(np.arange(n_persona, dtype=np.int64), n_field)
    f = np.tile(np.arange(n_field, dtype=np.int64), n_persona)
    b.const("persona-")
    b.number(p)
    b.const("/field-")
    b.number(f)
    return b.build()


def c5_pattern_sets(n_sets=64, seed=0xC5, n_persona=100000, n_field=100):
    """64 wildcard lists of 1-9 patterns in the shapes SURVEY.md §8(d) names."""
    rng = np.random.default_rng(seed)
    sets = []
    for _ in range(n_sets):
        pats = []
        for _ in range(int(rng.integers(1, 10))):
            shape = int(rng.integers(0, 6))
            p, f = int(rng.integers(0, n_persona)), int(rng.integers(0, n_field))
            if shape == 0:
                pats.append(f"persona-{p}/*")
            elif shape == 1:
                pats.append(f"persona-{p // 10}*/field-{f}")
            elif shape == 2:
                pats.append(f"*/field-{f}")
            elif shape == 3:
                pats.append(f"persona-{p}/field-{f}")
            elif shape == 4:
                pats.append(f"nomatch-{p}")
            else:
                pats.append("*" if rng.random() < 0.1 else f"*-{p % 1000}/*")
        sets.append(pats)
    return sets


# ---- C3: text_adventure-derived ---------------------------------------------------------------
def example_batches():
    """The reference's example programs as resolver batches: for each `examples/*.json5`, its default_state and the
    strings recursive_interpolate sends to the resolver for every top-level task, in order.  Derived from the files by
    the test tooling's `gen_golden.py examples` (load_program + the interp.rs:179-246 traversal) and committed as
    data/example_batches.json; the CPU test suite re-derives it from the reference tree when that is present."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "example_batches.json")) as f:
        return json.load(f)


_EX = example_batches()
C1_TEMPLATES = _EX["hello_world.json5"]["templates"]          # BASELINE.json config 1
C2_TASKS = _EX["math.json5"]["tasks"]                         # config 2: per task, the strings of its tree walk
C3_DEFAULT_INSERTS = _EX["text_adventure.json5"]["default_state"]["inserts"]  # text_adventure.json5:6-13
# config 3: every string of text_adventure.json5's top-level tasks + the nested-key forms of README.md:37
C3_TEMPLATES = _EX["text_adventure.json5"]["templates"] + ["{question-{i}}", "{persona_name}/answer-{i}"]


def c3_state(s, rng):
    st = dict(C3_DEFAULT_INSERTS)
    st.update({"i": (s % 16) + 1, "persona_name": f"P{s}", "stage": "first", "history_list": [], "history_text_llm": "",
               "scenario": "This is a text adventure game where you play as a knight errant number %d." % s})
    for q in range(1, 17):
        ln = int(rng.integers(20, 121))
        st[f"question-{q}"] = "".join(chr(c) for c in ALPHABET[rng.integers(0, len(ALPHABET), size=ln)])
    return st
