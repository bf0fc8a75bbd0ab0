#!/usr/bin/env python
"""Top stall-sample CUDA source lines of an ncu report: python profiles/hot_src.py <file.ncu-rep> [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = {}; fname = ""; hdr = None; cur = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; si = hdr.index("# Samples"); continue
    if not hdr or len(r) <= si: continue
    if r[0] != "":
        cur = (fname, r[0], r[1].strip()[:120]); agg.setdefault(cur, 0)
    elif cur is not None:
        try: agg[cur] += int(r[si] or 0)
        except ValueError: pass
tot = sum(agg.values())
print("total samples", tot)
for (f, ln, src), s in sorted(agg.items(), key=lambda kv: -kv[1])[:N]:
    print("%6d %5.1f%%  %s:%s  %s" % (s, 100.0 * s / max(tot, 1), f, ln, src))
