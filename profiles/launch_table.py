#!/usr/bin/env python
"""Print an ncu launch list (gpu__time_duration.sum CSV) in launch order or aggregated.
    python profiles/launch_table.py gpurun_out/launches_X.csv [--agg] [--from i --to j]"""
import csv, io, re, sys
from collections import OrderedDict
path = sys.argv[1]
agg = "--agg" in sys.argv
lo = int(sys.argv[sys.argv.index("--from") + 1]) if "--from" in sys.argv else 0
hi = int(sys.argv[sys.argv.index("--to") + 1]) if "--to" in sys.argv else 10**9
lines = [l for l in open(path, errors="ignore") if not l.startswith("==")]
rows = [r for r in csv.DictReader(io.StringIO("".join(lines))) if r["Metric Name"] == "gpu__time_duration.sum"]
def short(n):
    n = re.sub(r"\(.*", "", n); n = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", n)
    return n[:70]
def us(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
if agg:
    d = OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"]); d.setdefault(k, [0, 0.0]); d[k][0] += 1; d[k][1] += us(r)
    tot = sum(v[1] for v in d.values())
    for k, (n, t) in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %4d %9.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
    print("total %.1f us, %d launches" % (tot, len(rows)))
else:
    for i, r in enumerate(rows):
        if lo <= i < hi:
            print(i, short(r["Kernel Name"]), r.get("Grid Size"), "%.1f" % us(r))
