#!/usr/bin/env python
"""Reads the lines tools/ab_k1.py printed for two libraries and names the one to keep: the candidate (first argument) if
its k=1 rktvd1 fraction beats the baseline's (second argument) by more than 1.5 % and no other pair loses more than 1 %.
    python tools/ab_pick.py results.txt candidate.so baseline.so"""
import re
import sys

res = {}
for ln in open(sys.argv[1]):
    m = re.match(r"(\S+)\s+(?:\d+x\d+\s+)?k=(\d) rktvd(\d) \(\w+\): ([\d. ]+)\(", ln)
    if m:
        res[(m.group(1), int(m.group(2)), int(m.group(3)))] = max(float(v) for v in m.group(4).split())
cand, base = sys.argv[2], sys.argv[3]
pairs = sorted({(k, o) for (lib, k, o) in res if lib == cand} & {(k, o) for (lib, k, o) in res if lib == base})
ok = bool(pairs) and res[(cand, 1, 1)] > 1.015 * res[(base, 1, 1)] and all(res[(cand, k, o)] > 0.99 * res[(base, k, o)] for k, o in pairs)
print(cand if ok else base)
