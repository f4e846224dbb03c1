import sys, os
sys.path.insert(0, '/root/repo')
import interpolation_engine_b200 as ie
if len(sys.argv) > 1: ie.LIB_PATH = os.path.join(os.path.dirname(ie.LIB_PATH), sys.argv[1])
eng = ie.Engine(0)
ins = {"k%d" % k: "value-%d" % k for k in range(40)}
giant = "".join("some literal text %03d {k%d} " % (k, k % 40) for k in range(600))
templates = ["plain {k1}", giant, "{k2}{k3}", giant[:4000], "x"] + ["t%d {k%d}" % (k, k % 40) for k in range(300)]
table = eng.pack(ie.PackedInserts.from_dict(ins))
best = min(eng.resolve_batch(table, templates).kernel_ms for _ in range(6))
print(sys.argv[1:], 'giant batch kernel_ms', best)
# a batch where every 20th template has a 20-byte key, 50k templates
import random
ins2 = {("key-%02d" % k): "v%d" % k for k in range(50)}; ins2["a_key_of_twenty_bytes"] = "LONG"
t2 = [("t%d {key-%02d} {a_key_of_twenty_bytes}" % (k, k % 50)) if k % 20 == 0 else ("t%d {key-%02d} and {key-%02d}" % (k, k % 50, (k * 7) % 50)) for k in range(200000)]
tab2 = eng.pack(ie.PackedInserts.from_dict(ins2))
best2 = min(eng.resolve_batch(tab2, t2).kernel_ms for _ in range(6))
print(sys.argv[1:], '200k templates, 5 % with a 20-byte key: kernel_ms', best2)
