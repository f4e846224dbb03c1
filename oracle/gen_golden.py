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


# ---- replace_map / goto_map: the twin's execute_task drives them -------------------------
# interpolation_engine.py:1689-1770 holds both commands inlined in execute_task; runtime.rs:1085-1133, 1649-1752 is the
# Rust side.  The two differ outside the subset generated here (documented where each restriction is applied):
#   * Rust resolves goto_map entries lazily and in order, Python resolves every key and value up front
#     -> every key / value template of a generated case resolves;
#   * Python's `str()` of a list / bool / float is its repr, Rust's value_to_string concatenates / lower-cases
#     -> inserts are strings and integers only;
#   * a replace_map item that IS one simple key is looked up typed by Python and rescanned (`recursive_replace`), Rust
#     resolves it once -> such items resolve to brace-free strings;
#   * on a resolver error Python falls back to the literal 'NULL' map for the whole item, Rust only when the item is a
#     simple key that fails (runtime.rs:1705-1711; every other error propagates through `?`)
#     -> failing cases with a NULL map use simple-key items; without a NULL map only "it is an error" is compared;
#   * Python's `$` also matches before a trailing newline -> cases where a tested text ends in a newline and the
#     pattern does not end in '*' are dropped (detected by wrapping the twin's matcher).
MAPS_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "python_twin_maps.json")
EXAMPLES = "/root/reference/examples"


def example_map_tasks():
    """Every replace_map / goto_map task of the reference's example programs, loaded through the oracle's
    restatement of parser.rs (the twin's own loader needs the json5 package)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tests import oracle_lib
    orc = oracle_lib.load()
    found = []

    def walk(t, src):
        if isinstance(t, dict):
            if t.get("cmd") in ("replace_map", "goto_map"):
                found.append((src, t))
            for v in t.values():
                walk(v, src)
        elif isinstance(t, list):
            for v in t:
                walk(v, src)

    for name in sorted(os.listdir(EXAMPLES)):
        if name.endswith(".json5"):
            kind, prog = orc.call("load_program", text=open(os.path.join(EXAMPLES, name), encoding="utf-8").read())
            assert kind == "ok", (name, prog)
            walk(prog, name)
    return found


def template_keys(twin, text):
    """Names a template interpolates (top level and nested), by brute force over the text."""
    return set(re.findall(r"\{([^{}]*)\}", text))


def gen_maps():
    import asyncio
    twin = load_twin()
    twin.log_sink = open(os.devnull, "w")
    rng = random.Random(0x3A95)
    unsafe = []
    real_match = twin.is_wildcard_match

    def watched(p, s):
        if s.endswith("\n") and not p.endswith("*"):
            unsafe.append((p, s))
        return real_match(p, s)

    twin.is_wildcard_match = watched
    signal.signal(signal.SIGALRM, _alarm)

    def run(task, ins):
        """-> {"ok": value} | {"err": True} | None (dropped: outside the PY == RS subset, or does not terminate)"""
        del unsafe[:]
        st = {"inserts": json.loads(json.dumps(ins))}
        t = dict(task, traceback_label="t")
        signal.setitimer(signal.ITIMER_REAL, 0.5)
        try:
            r = asyncio.run(twin.execute_task(st, t, {}, {}, "x"))
            if task["cmd"] == "goto_map":
                res = {"ok": r["goto_target"] if r else "CONTINUE"}
            else:
                res = {"ok": st["inserts"][task["output_name"]]}
        except (twin.InterpolationException, AssertionError):
            res = {"err": True}
        except (RecursionError, _Timeout, MemoryError):
            return None
        finally:
            signal.setitimer(signal.ITIMER_REAL, 0)
        return None if unsafe else res

    words = ["look", "go north", "Hello   world", "a", "", "x  y", "the  door\n\n\n\nopens", " pad ", "1", "2", "3", "(none)",
             "(unset)", "/undo", "/restart", "/summarize", "(peek)", "true", "false", "0", "first", "action", "query", "undo", "other",
             "Morning", "Noon", "Evening", "Night", "Benjamin"]
    tags = ["first-output", "action-output", "query-output", "query", "action"]

    def llm_text():
        parts = []
        for _ in range(rng.randint(0, 5)):
            r = rng.random()
            if r < 0.5:
                tg = rng.choice(tags)
                parts.append("<%s>%s</%s>" % (tg, rng.choice(words), tg))
            elif r < 0.75:
                parts.append(rng.choice(["  ", " ", "\n", "\n\n\n", "\n\n\n\n", "   "]))
            else:
                parts.append(rng.choice(words))
        return "".join(parts)

    cases = []

    def add(src, task, ins, loose=False):
        r = run(task, ins)
        if r is None:
            return
        c = {"src": src, "fn": task["cmd"], "inserts": ins, "py": r}
        if loose:
            c["loose"] = True  # both sides fail, with different messages: only "it is an error" is pinned
        if task["cmd"] == "goto_map":
            c["args"] = {"text": task["text"], "target_maps": task["target_maps"]}
        else:
            c["args"] = {"item": task["item"], "wildcard_maps": task["wildcard_maps"], "repeat_until_done": bool(task.get("repeat_until_done", False))}
        cases.append(c)

    # 1. the example programs' own maps against synthetic states
    for src, task in example_map_tasks():
        texts = [task["text"]] if task["cmd"] == "goto_map" else [task["item"]]
        maps = task["target_maps"] if task["cmd"] == "goto_map" else task["wildcard_maps"]
        for m in maps:
            for k, v in m.items():
                texts += [k] + ([v] if isinstance(v, str) else [])
        keys = sorted(set().union(*[template_keys(twin, t) for t in texts]) - {"1", "2", "3", "4", "5", "6"})
        item_keys = template_keys(twin, texts[0])
        has_null = any("NULL" in m for m in maps)
        simple_item = twin.get_simple_insertkey(texts[0]) is not None
        for trial in range(40):
            ins = {}
            for k in keys:
                if k == "history_text_base":
                    ins[k] = llm_text()
                elif simple_item and k in item_keys:
                    ins[k] = rng.choice(words)  # a simple-key item that resolves to a number stays a number in Python, Rust renders it
                else:
                    ins[k] = rng.choice(words + [7, 12])
            if trial % 8 == 7 and item_keys:  # the text / item does not resolve
                del ins[sorted(item_keys)[0]]
                if task["cmd"] == "replace_map" and not simple_item and has_null:
                    continue  # Python: NULL map; Rust: the error propagates
                add(src + ":%d" % task.get("line", 0), task, ins, loose=not has_null)
            else:
                add(src + ":%d" % task.get("line", 0), task, ins)

    # 2. random maps over a small alphabet
    lits = ["a", "b", "ab", "-", "/", " ", "x", "", "|"]

    def rand_pattern():
        return "".join(rng.choice(lits + ["*", "*", "{k}", "{n}"]) for _ in range(rng.randint(0, 5)))

    def rand_value(ncap):
        # capture references only up to the pattern's star count: without a star Python's findall yields the whole
        # match as {1}, Rust's captures are empty
        caps = ["{%d}" % rng.randint(1, ncap)] if ncap else []
        return "".join(rng.choice(lits + caps + ["{k}", "{n}", "[", "]"]) for _ in range(rng.randint(0, 5)))

    for trial in range(1500):
        ins = {"k": rng.choice(["a", "b", "ab", "a-b", "", "x/x"]), "n": rng.randint(0, 12), "t": "".join(rng.choice("ab-/ x|") for _ in range(rng.randint(0, 8)))}
        maps = []
        for _ in range(rng.randint(1, 5)):
            p = rand_pattern()
            ncap = p.count("*")
            maps.append({p: rand_value(ncap)})
        if rng.random() < 0.3:
            maps.insert(rng.randint(0, len(maps)), {"NULL": rng.choice(["fallback", "", "n{n}"])})
        if rng.random() < 0.5:
            item = rng.choice(["{t}", "{t}", "{k}", "{missing}", "{{q}}"])  # a simple key (resolves to a brace-free string, or fails)
            ins["q"] = "k"
        else:
            item = "".join(rng.choice(lits + ["{k}", "{n}", "{t}"]) for _ in range(rng.randint(1, 5)))
            if twin.get_simple_insertkey(item):
                item += "."
        fails = item == "{missing}"
        has_null = any("NULL" in m for m in maps)
        # entries that cannot resolve (a capture the pattern does not have) are reached only sometimes; with a NULL map
        # Python would fall back where Rust propagates, so such cases are kept only without one
        shape = rng.random()
        if shape >= 0.8:  # goto_map has no captures
            maps = [{k: re.sub(r"\{\d\}", "", v)} for m in maps for k, v in m.items()]
        if shape < 0.6:
            task = {"cmd": "replace_map", "item": item, "output_name": "out", "wildcard_maps": maps, "repeat_until_done": rng.random() < 0.5}
        elif shape < 0.8:
            if fails:
                continue
            it = rng.choice([[item, 5, None, "b-a"], {"key " + item: item, "z": [item]}])
            task = {"cmd": "replace_map", "item": it, "output_name": "out", "wildcard_maps": maps, "repeat_until_done": False}
        else:
            tm = [{k: "T" + v} for m in maps for k, v in m.items()]
            task = {"cmd": "goto_map", "text": item, "target_maps": tm}
        r = run(task, ins)
        if r is None:
            continue
        if task["cmd"] == "goto_map" and any(not resolves(twin, ins, s) for m in task["target_maps"] for kv in m.items() for s in kv):
            continue  # Python resolves every entry up front, Rust only those it reaches
        if "err" in r:
            if has_null and not (task["cmd"] == "goto_map"):
                continue
            add("fuzz", task, ins, loose=True)
        else:
            add("fuzz", task, ins)
    with open(MAPS_OUT, "w") as f:
        json.dump({"cases": cases}, f, ensure_ascii=True, indent=0)
    print({"maps_cases": len(cases), "errors": sum("err" in c["py"] for c in cases)})


# ---- the example programs as resolver batches (BASELINE.json configs 1-3) ---------------------------
# Each example is loaded through the restatement of parser.rs (add_line_numbers + load_program) and every top-level task
# of its `order` is walked the way recursive_interpolate walks it: the strings that reach interpolate_inserts, in order.
# The fixture holds those derived batches (not the program text); tests/test_oracle_golden.py re-derives them from
# /root/reference when the tree is present and compares.
EXAMPLES_OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "interpolation_engine_b200", "data", "example_batches.json")


def example_batches():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from tests import oracle_lib
    orc = oracle_lib.load()
    out = {}
    for name in sorted(os.listdir(EXAMPLES)):
        if not name.endswith(".json5"):
            continue
        kind, prog = orc.call("load_program", text=open(os.path.join(EXAMPLES, name), encoding="utf-8").read())
        assert kind == "ok", (name, prog)
        tasks = []
        for task in prog["order"]:
            kind, tr = orc.call("interpolation_trace", value=task)
            assert kind == "ok"
            tasks.append({"cmd": task.get("cmd"), "line": task.get("line"), "templates": tr["templates"], "lookups": tr["lookups"]})
        out[name] = {"default_state": prog["default_state"], "tasks": tasks,
                     "templates": [t for task in tasks for t in task["templates"]]}
    return out


def gen_examples():
    out = example_batches()
    os.makedirs(os.path.dirname(EXAMPLES_OUT), exist_ok=True)
    with open(EXAMPLES_OUT, "w") as f:
        json.dump(out, f, ensure_ascii=True, indent=0, sort_keys=True)
    print({k: len(v["templates"]) for k, v in out.items()})


def resolves(twin, ins, s):
    try:
        twin.interpolate_inserts(ins, s)
        return True
    except Exception:
        return False


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "twin"):
        main()
    if which in ("all", "maps"):
        gen_maps()
    if which in ("all", "examples"):
        gen_examples()
