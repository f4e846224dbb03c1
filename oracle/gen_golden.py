"""Generate tests/golden/python_twin.json by executing the reference's own Python resolver.

ORACLE TOOLING — test infrastructure only.  Run in the build container (needs /root/reference):

    python oracle/gen_golden.py

The reference's Rust resolver cannot be built here, and the reference ships no tests, so the only
executable form of the reference on this path is its Python twin
(/root/reference/src/interpolation_engine/interpolation_engine.py:426-567, 1436-1494).  It imports
once `json5` and `prompt_toolkit` are stubbed.  Vectors are restricted to the
subset where Python == Rust: error *kinds* and keys rather than message
texts; the replaying test skips lists/floats spliced into text and `}{`-style inputs (Rust panics,
Python returns).  The committed JSON is what the CPU tests replay; nothing reads /root/reference
at test time.
"""
import importlib
import json
import os
import random
import re
import signal
import sys
import types

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "python_twin.json")


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"from_dict": staticmethod(lambda *a, **k: None),
                               "__init__": lambda self, *a, **k: None})


def load_twin():
    for name in ["json5", "prompt_toolkit"] + ["prompt_toolkit." + m for m in (
            "application", "filters", "history", "key_binding", "layout", "buffer", "document",
            "layout.dimension", "layout.controls", "styles", "widgets", "data_structures",
            "formatted_text", "layout.containers", "layout.layout", "keys", "shortcuts", "patch_stdout")]:
        sys.modules.setdefault(name, _Dummy(name))
    sys.path.insert(0, REF_SRC)
    return importlib.import_module("interpolation_engine.interpolation_engine")


def classify_error(e):
    msg = str(e)
    if msg.startswith("Argument interpolation key '"):
        return {"kind": "arg", "key": msg.split("'")[1]}
    if msg.startswith("Tried to interpolate empty string"):
        return {"kind": "empty", "key": ""}
    if msg.startswith("Could not find variable '"):
        m = re.match(r"Could not find variable '(.*)' in interpdata", msg, re.S)
        return {"kind": "not_found", "key": m.group(1)}
    if msg.startswith("Error: The following content has"):
        m = re.match(r'Error: The following content has \d+ \'\{\' and \d+ \'\}\':\n\n"""(.*)\n"""$', msg, re.S)
        return {"kind": "uneven", "key": m.group(1)}
    if msg.startswith("Error: trying to interpolate variable '"):
        m = re.match(r"Error: trying to interpolate variable '(.*)' of type", msg, re.S)
        return {"kind": "unsupported", "key": m.group(1)}
    raise RuntimeError("unclassified python error: " + msg)


BS = "\\"
BASE_INSERTS = {
    "i": 3, "question-3": "Q3?", "name": "tom", "x": "{y}", "y": "z", "z": "ZED", "k": "i", "result": 3,
    "ARG1": "a" + BS + "{b" + BS + "}", "slot-7": 12, "idx-12": 40, "q-40": "deep", "e": "",
    "lst": ["x", "y"], "b": True, "persona_name": "Benjamin", "u": "foo " + BS + "{bar." + BS + "}",
    "obj": {"a": 1}, "nul": None, "dot": "end.", "esc": BS + "}tail", "bs": "trail" + BS,
    "open": "{", "close": "}", "pair": "{name}", "sent": "lit .〠 x", "tnes": "〠.y",
}

HAND_TEMPLATES = [
    "Hello, world!", "The result is {result}!\n", "{question-{i}}", "{persona_name}/answer-{i}",
    "{q-{idx-{slot-7}}}", "a {q-{idx-{slot-7}}} b {name}", "{result}", "{lst}", "{x}", "-{x}-", "{{x}}",
    "x{{x}}", "{{k}}", "{question-{{k}}}", "keep " + BS + "{name" + BS + "}", "{ARG1}", "arg={ARG1}", "{ARG2}",
    "arg={ARG2}", "{ARG}", "{}", "", "{missing1} {missing2}", "{a}}", "{a", "a}", BS + BS + "{name}",
    BS + "{name}", "lit .〠 x", "〠. x", "{e}", "[{e}]", "b={b}", "{b}", "{obj}", "o={obj}", "{nul}",
    "n={nul}", "foo " + BS + "{bar." + BS + "}", "a." + BS + "}", "hi {name}." + BS + "}", "{u}", "v={u}", BS + "{.",
    "{dot}" + BS + "}", "." + BS + "}" + BS + "}", "x." + BS + "}." + BS + "}", "{esc}", ".{esc}", "{dot}{esc}",
    "{name} {name} {name}", "{ name }", "{name }", "{{{k}}}", "{{{{k}}}}", "{pair}", "p={pair}", "{open}", "o={open}x",
    "{close}", "c={close}", "{bs}", "{bs}{name}", "x{bs}" + BS + "{", "{sent}", "s={sent}", "{tnes}", ".{tnes}",
    "{k}{i}", "{question-{k}}", "{question-{{k}}} and {q-{idx-{slot-7}}}", "{{missing}}", "{q-{missing}}",
    "t{q-{missing}} {alsomissing}", "a\nb{name}\n", "{3}", "{{result}}", "x{{result}}", "{HH:MMx}",
    "{na" + BS + "{me}", "{na" + BS + "}me}", "a" + BS, BS, BS + BS, "{" + BS + "}", "{name" + BS + "}",
]

WILDCARDS = [
    ("Benjamin/*", "Benjamin/"), ("Benjamin/*", "benjamin/x"), ("*", ""), ("", ""), ("", "a"),
    ("a*b*c", "a--b--c--b--c"), ("a.c", "abc"), ("a.c", "a.c"), ("enable_*", "enable_suggestions"),
    ("enable_*", "enable"), ("(*)", "(look)"), ("*|*", "a|b\nc"), ("*|*", "x|y|z"), ("**", "abc"),
    ("a**b", "ab"), ("a*", "a"), ("*a", "a"), ("*a", "ba"), ("*a", "ab"), ("a*a", "a"), ("a*a", "aa"),
    ("persona-12/*", "persona-12/field-3"), ("persona-1*/field-3", "persona-12/field-3"),
    ("*/field-3", "persona-12/field-33"), ("[a]+?^$", "[a]+?^$"), ("\\d", "5"), ("*  *", "a  b"),
    ("*\n\n\n*", "x\n\n\ny"), ("*<query>*</query>*", "pre<query>look</query>post"), ("a*b", "a\nb"),
]


def fuzz_cases(rng, n):
    alphabet = ["{", "}", "{", "}", BS, ".", "a", "b", "-", " ", "〠", "n1", "k1", "k2", "k3", "miss"]
    keys = ["a", "b", "k1", "k2", "k3", "a-b", "b-a", "aa", "ab", "n1", "a.", ".a", "k1-v"]
    for _ in range(n):
        ins = {}
        for k in rng.sample(keys, rng.randint(2, len(keys))):
            r = rng.random()
            if r < 0.15:
                ins[k] = rng.randint(0, 99)
            elif r < 0.75:
                ins[k] = rng.choice(["a", "b", "k1", "k2", "v", "", "a-b", "x y", "end.", "{a}", "{k1}", BS + "{q" + BS + "}",
                                     "." + BS + "}", "d." + BS + "}", BS + "}", "t" + BS, "}", "{", "a}b", "{k2}-{k3}",
                                     "〠.", ".〠"])
            else:
                ins[k] = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 6)))
        t = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 14)))
        if rng.random() < 0.5:  # bias toward well-formed groups
            parts = []
            for _ in range(rng.randint(1, 4)):
                k = rng.choice(keys + ["miss", "{k1}", "a-{k2}", "{k3}-b", "{{k1}}"])
                parts.append(rng.choice(["", "x", ". ", BS + "{", BS + "}", "." + BS + "}"]) + "{" + k + "}")
            t = "".join(parts) + rng.choice(["", ".", BS + "}", "tail"])
        yield ins, t


class _Timeout(Exception):
    pass


def _alarm(*_):
    raise _Timeout


def main():
    twin = load_twin()
    rng = random.Random(0x60DE)
    out = {"interpolate": [], "simple_key": [], "escape": [], "unescape": [], "wildcard": [], "captures": []}
    signal.signal(signal.SIGALRM, _alarm)

    def run_interp(ins, t):
        # the twin has no bound on self-referential values either: cut those cases off
        signal.setitimer(signal.ITIMER_REAL, 0.25)
        try:
            r = twin.interpolate_inserts(ins, t)
            return {"ok": r}
        except (twin.InterpolationException, AssertionError) as e:
            return {"err": classify_error(e)}
        except (RecursionError, _Timeout, MemoryError):
            return None
        finally:
            signal.setitimer(signal.ITIMER_REAL, 0)

    for t in HAND_TEMPLATES:
        r = run_interp(BASE_INSERTS, t)
        if r is not None:
            out["interpolate"].append({"inserts": "base", "template": t, "py": r})
    for ins, t in fuzz_cases(rng, 6000):
        r = run_interp(ins, t)
        if r is not None:
            out["interpolate"].append({"inserts": ins, "template": t, "py": r})
    for t in HAND_TEMPLATES + ["{a}{b}", "{{a}{b}}", "{a{b}", "ab", "a", "{", "}", "{}", "{{}}", "x{a}", "{a}x", "{a}}"]:
        k = twin.get_simple_insertkey(t)
        out["simple_key"].append({"content": t, "py": k if k else None})  # '' is falsy in the twin's callers
    vals = ["{x}", "a{b}c", BS + "{x" + BS + "}", {"k{": ["a}", 1, None, True]}, {"k" + BS + "{": ["a" + BS + "}", 1]},
            ["{", "}"], 5, "plain"]
    for v in vals:
        out["escape"].append({"value": v, "py": twin.recursive_escape(v)})
        out["unescape"].append({"value": v, "py": twin.recursive_unescape(v)})
    for p, s in WILDCARDS:
        out["wildcard"].append({"pattern": p, "text": s, "py": twin.is_wildcard_match(p, s)})
        if "*" in p and twin.is_wildcard_match(p, s):
            out["captures"].append({"pattern": p, "text": s, "py": twin.get_wildcard_matches(p, s)})
    wr = random.Random(7)
    for _ in range(400):
        p = "".join(wr.choice("ab*/-") for _ in range(wr.randint(0, 6)))
        s = "".join(wr.choice("ab/-\n") for _ in range(wr.randint(0, 8)))
        if s.endswith("\n") and not p.endswith("*"):
            continue  # Python's `$` also matches before a trailing newline; Rust's does not
        m = twin.is_wildcard_match(p, s)
        out["wildcard"].append({"pattern": p, "text": s, "py": m})
        if "*" in p and m:
            out["captures"].append({"pattern": p, "text": s, "py": twin.get_wildcard_matches(p, s)})
    out["base_inserts"] = BASE_INSERTS
    with open(OUT, "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
