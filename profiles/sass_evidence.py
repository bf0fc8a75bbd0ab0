#!/usr/bin/env python
"""Blackwell-specific SASS mnemonics per kernel of the built library (run where the .so was built, no GPU needed):
    python profiles/sass_evidence.py > profiles/r02_sass_blackwell.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "canonicalsg2im_b200", "libcsg2im.so")
KEYS = ("UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
fn = None
per = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        if op.split(".")[0] in KEYS:
            per[fn][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print("# SASS evidence (cuobjdump -sass canonicalsg2im_b200/libcsg2im.so, sm_100a), round 2, final state")
print("#")
print("# Per kernel: counts of the Blackwell-specific mnemonics -- tcgen05.mma = UTCHMMA (.2CTA = cta_group::2),")
print("# tcgen05.commit = UTCBAR, tcgen05.ld = LDTM, tcgen05.alloc = UTCATOMSWS, TMA tile / tile::gather4 loads = UTMALDG,")
print("# cp.async.bulk = UBLKCP, mbarrier = SYNCS -- and the distinct instruction forms.  Made by profiles/sass_evidence.py.")
print("#")
tot = collections.Counter()
for c in per.values():
    for op, n in c.items():
        tot[op.split(".")[0]] += n
print("\nTOTAL over the library: " + ", ".join("%s %d" % (k, tot[k]) for k in sorted(tot)))
for (fn, c), name in zip(per.items(), names):
    base = collections.Counter()
    for op, n in c.items():
        base[op.split(".")[0]] += n
    if not (base.keys() - {"SYNCS"}):
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(.*", "", name)
    print("\n## " + name)
    print("counts: " + ", ".join("%s %d" % (k, base[k]) for k in sorted(base)))
    for op, n in sorted(c.items(), key=lambda kv: (kv[1], kv[0])):
        if not op.startswith("SYNCS"):
            print("%8d  %s" % (n, op))
