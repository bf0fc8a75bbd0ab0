#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/<tag>_summary.md (run here, no GPU needed).

    python profiles/summarize.py r01fp32
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    return name.strip()


def launches(tag):
    path = os.path.join(OUT, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path, errors="ignore") if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v
    return agg


def raw(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return {}
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        d[h] = (v, u)
    return d


def main():
    tag = sys.argv[1]
    out = ["# ncu summary `%s`" % tag, "",
           "Per-launch times below come from `ncu --metrics gpu__time_duration.sum --clock-control none` over "
           "`python bench.py --steps 1 --warmup 3 --profile` (one profiled step; cold-cache, serialised: compare SHARES).", ""]
    agg = launches(tag)
    if agg:
        total = sum(v[1] for v in agg.values())
        out += ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
        for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
            out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / total))
        out += ["", "total device time in list: %.1f us over %d launches" % (total, sum(v[0] for v in agg.values())), ""]
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_%s_*.ncu-rep" % tag))):
        d = raw(rep)
        if not d:
            continue
        out += ["## `%s` (`ncu --set full`, one launch)" % os.path.basename(rep), "", "| metric | value | unit |", "|---|---:|---|"]
        kn = d.get("Kernel Name", ("?", ""))[0]
        out.append("| kernel | `%s` | |" % short(kn))
        for k in KEYS:
            if k in d:
                out.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
        out.append("")
    path = os.path.join(ROOT, "profiles", "%s_summary.md" % tag)
    open(path, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
