"""Latency of the JSON-level single calls (ie_call): what one replace_map iteration costs and where it goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
eng = ie.Engine(0)
ins = {"name": "Ada", "i": 3, "question-3": "why?", "x": "a b"}
ins_big = dict(ins); ins_big.update({"k%d" % i: "value %d" % i for i in range(5000)})


def t(label, fn, reps, **kw):
    eng.call(fn, **kw)
    t0 = time.perf_counter()
    for _ in range(reps): r = eng.call(fn, **kw)
    dt = (time.perf_counter() - t0) / reps
    print("%-60s %9.1f us  %s" % (label, dt * 1e6, repr(r)[:60]), flush=True)
    return dt

t("interpolate_inserts, 4 inserts", "interpolate_inserts", 200, inserts=ins, content="hi {name} {question-{i}}")
t("interpolate_inserts, 5004 inserts", "interpolate_inserts", 50, inserts=ins_big, content="hi {name} {question-{i}}")
t("wildcard_captures", "wildcard_captures", 200, pattern="*-*", text="a-b-c")
maps = [{"<t>*</t>*": "{1}|{2}"}, {"* *": "{2}_{1}"}, {"zzz": "never"}]
t("replace_map, no match, 3 patterns", "replace_map", 200, inserts=ins, item="plain", wildcard_maps=maps, repeat_until_done=False)
t("replace_map, one match + captures", "replace_map", 200, inserts=ins, item="a b", wildcard_maps=maps, repeat_until_done=False)
t("replace_map, same with 5004 inserts", "replace_map", 30, inserts=ins_big, item="a b", wildcard_maps=maps, repeat_until_done=False)
grow = [{"*": "z{1} "}]
for n_it in ():
    dt = t("replace_map, repeat_until_done, text grows 2 B per iteration (limit)", "replace_map", 1, inserts=ins, item="s", wildcard_maps=grow, repeat_until_done=True)
