#!/usr/bin/env python
"""SASS opcode histogram of every kernel in cmusphinx_b200/libb200sphinx.so (cuobjdump -sass), the
evidence of what the hand-written kernels compile to: UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,
UBLKCP = cp.async.bulk (bulk TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.
    python tools/sass_histogram.py > profiles/r2_sass_histogram.json"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "cmusphinx_b200", "libb200sphinx.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"b200::\(anonymous namespace\)::|\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        cur = kern.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "FFMA", "FADD", "FMNMX3", "FMNMX", "VIMNMX", "VIMNMX3", "F2I", "LOP3",
       "IMAD", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "SHFL", "VOTE", "DFMA", "DMUL", "DADD", "BAR"]
res = {}
for k, c in kern.items():
    res[k] = {"instructions": sum(c.values()), **{o: c[o] for o in KEY if c[o]}}
tot = collections.Counter()
for c in kern.values():
    tot.update(c)
print(json.dumps({"library": "cmusphinx_b200/libb200sphinx.so", "kernels": len(res),
                  "totals": {o: tot[o] for o in KEY if tot[o]}, "per_kernel": res}, indent=1))
