"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of OUR kernels (last pass of the
script = second half of the launches) and the longest individual launches.  usage: launch_summary.py file.csv [n_passes]"""
import collections
import csv
import re
import sys

path, passes = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r.get("Grid Size", "")) for r in csv.DictReader(lines)]
ours = [x for x in rows if not x[0].startswith(("void at::", "at::", "void c10", "void at_cuda"))]
last = ours[len(ours) - len(ours) // passes:]
agg = collections.defaultdict(lambda: [0, 0.0])
for n, t, g in last:
    k = re.sub(r"\(.*", "", n)[:80]
    agg[k][0] += 1
    agg[k][1] += t
tot = sum(v[1] for v in agg.values())
print(f"{path}: {len(rows)} launches, {len(ours)} ours, last pass {len(last)} launches, {tot / 1e6:.1f} ms")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e6:9.1f} ms {v[0]:4d}  {100 * v[1] / tot:5.1f}%  {k}")
print("longest launches:")
for n, t, g in sorted(last, key=lambda x: -x[1])[:12]:
    print(f"{t / 1e6:8.2f} ms {g:>16} {re.sub(r'[(].*', '', n)[:70]}")
