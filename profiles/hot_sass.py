#!/usr/bin/env python
"""Top stall-sample SASS lines of an ncu report: python profiles/hot_sass.py <file.ncu-rep> [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; body = rows[2:]
si = hdr.index("# Samples") if "# Samples" in hdr else hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
idx = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    print("%5d %6s %5.1f%%  %s" % (i, r[si], 100.0 * int(r[si] or 0) / max(tot, 1), r[1].strip()[:110]))
