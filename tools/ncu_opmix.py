#!/usr/bin/env python
"""Executed SASS instruction mix per cell of the first kernel in an .ncu-rep (source page).
usage: python tools/ncu_opmix.py rep.ncu-rep <cells>"""
import csv
import re
import subprocess
import sys
from collections import Counter

rep, cells = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
ks = []
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        ks.append({"name": r[1], "hdr": None, "rows": []})
    elif r and r[0] == "Address":
        ks[-1]["hdr"] = r
    elif ks and r:
        ks[-1]["rows"].append(r)
K = ks[0]
h = K["hdr"]
ie, src, si = h.index("Instructions Executed"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
c, st = Counter(), Counter()
for r in K["rows"]:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[src].strip()).split()[0].split(".")[0]
    c[op] += int(r[ie])
    st[op] += int(r[si])
tot = sum(c.values())
fp64 = sum(n for op, n in c.items() if op in ("DFMA", "DMUL", "DADD", "DSETP"))
print(K["name"])
print(f"thread-instructions per cell: total {tot * 32 / cells:.1f}  fp64-pipe {fp64 * 32 / cells:.1f}  other {(tot - fp64) * 32 / cells:.1f}")
for op, n in c.most_common(26):
    print(f"  {op:10s} {n * 32 / cells:8.2f}   stall samples {st[op]}")
