"""Top stall-sample SASS lines per kernel from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu -i rep --page source --csv | python tools/ncu_source_top.py [kernel_index] [top_n]"""
import csv
import sys

which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(sys.stdin))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
s, e = starts[which], starts[which + 1]
print("kernel:", rows[s][1][:100])
h = rows[s + 1]
ci = {n: i for i, n in enumerate(h)}
body = [r for r in rows[s + 2:e] if len(r) >= len(h)]
samp = ci["# Samples"]
tot = sum(int(r[samp]) for r in body)
print("total samples", tot, "instructions", len(body))
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(int(r[ci[n]]) for r in body) for n in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
order = sorted(range(len(body)), key=lambda i: -int(body[i][samp]))[:topn]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ci[n]]), n) for n in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {int(r[samp]):6d} {100 * int(r[samp]) / tot:5.1f}%  exec={r[ci['Instructions Executed']]:>8s}  {r[ci['Source']].strip()[:70]:70s} {top}")
