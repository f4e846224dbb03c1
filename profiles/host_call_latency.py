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
# a text of kilobytes against the 12 replace_map patterns of examples/text_adventure.json5:33-59 (shapes only)
long_text = ("The guard looks at you. " * 40) + "<q>What brings you here?</q>" + (" He waits." * 150)
ta_maps = [{"*<%s>*</%s>*" % (t, t): "{1}{3}"} for t in ("think", "aside", "ooc", "meta", "note", "plan")] + \
          [{"*[[*]]*": "{1}{3}"}, {"* \n*": "{1}\n{2}"}, {"*\n\n\n*": "{1}\n\n{2}"}, {"*  *": "{1} {2}"}, {"\n*": "{1}"}, {"* ": "{1}"}]
t("replace_map, %d B text, 12 patterns, none matches" % len(long_text), "replace_map", 100, inserts=ins, item=long_text, wildcard_maps=ta_maps[:7], repeat_until_done=False)
t("replace_map, same text, repeat_until_done (strips 150 double spaces..)", "replace_map", 5, inserts=ins, item=long_text.replace(". He", ".  He"), wildcard_maps=ta_maps, repeat_until_done=True)
t("goto_map, same text, 12 targets", "goto_map", 100, inserts=ins, text=long_text, target_maps=[{list(m)[0]: "@t%d" % i} for i, m in enumerate(ta_maps[:7])] + [{"*waits.": "@end"}])
