"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/prof_net_call.py: per-kernel totals of the LAST
network call.  usage: python tools/launch_summary.py gpurun_out/launches.csv [n_calls]"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1])))
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [(r[kn], float(r[mv].replace(",", ""))) for r in rows[hi + 1:] if len(r) >= len(h) and r[mv].replace(",", "").replace(".", "").isdigit()]
per_call = len(data) // n_calls
last = data[-per_call:]
agg = OrderedDict()
for name, ns in last:
    name = name.split("(")[0].replace("void ", "").replace("dexb::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values())
print(f"# {per_call} launches per network call, {tot / 1e6:.3f} ms (cold-cache, serialised: compare shares)")
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:44s} n={n:3d} {ns / 1e3:9.1f} us  {100 * ns / tot:5.1f}%")
