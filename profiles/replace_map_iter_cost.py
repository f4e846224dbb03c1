"""Why one replace_map iteration cost 10 ms in fuzz seed 30008: per-call time against the number of inserts and the text."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interpolation_engine_b200 as ie
from tests.fuzz_campaign import batch, CLOCK
eng = ie.Engine(0)
ins_all, _ = batch(30008)
maps = [{'{ARG1}**': 'z{1} '}, {'* a': '{2}{2}{3}'}, {'*': '{3}'}, {'a{a.}a k3\\m': 'z{2}'}]
item = '〠}\\k2\\ 〠-}k2'
print(len(ins_all), "inserts;", {k: (v if not isinstance(v, str) or len(v) < 40 else v[:40] + "...") for k, v in list(ins_all.items())[:30]})

def t(label, reps, fn, **kw):
    eng.call(fn, clock=CLOCK, **kw)
    t0 = time.perf_counter()
    for _ in range(reps): r = eng.call(fn, clock=CLOCK, **kw)
    print("%-64s %9.1f us  %s" % (label, (time.perf_counter() - t0) / reps * 1e6, repr(r)[:70]), flush=True)

for n in (2, 8, len(ins_all)):
    ins = dict(list(ins_all.items())[:n]); ins.setdefault("ARG1", ""); ins.setdefault("a.", "x")
    t("replace_map once, %d inserts" % len(ins), 50, "replace_map", inserts=ins, item=item, wildcard_maps=maps, repeat_until_done=False)
    t("  interpolate_inserts(item)", 50, "interpolate_inserts", inserts=ins, content=item)
    t("  interpolate_inserts('z{1} ') ", 50, "interpolate_inserts", inserts=dict(ins, **{"1": item}), content="z{1} ")
    t("  interpolate_inserts('plain') ", 50, "interpolate_inserts", inserts=ins, content="plain")
    t("  wildcard_captures('**', item)", 50, "wildcard_captures", pattern="**", text=item)
for k, v in ins_all.items():
    ins = {"ARG1": "", "a.": "x", k: v}
    t0 = time.perf_counter(); eng.call("interpolate_inserts", clock=CLOCK, inserts=ins, content=item); eng.call("interpolate_inserts", clock=CLOCK, inserts=ins, content=item); dt = (time.perf_counter() - t0) / 2
    if dt > 1e-3: print("slow with insert", repr(k), repr(v)[:80], "%.1f ms" % (dt * 1e3))
