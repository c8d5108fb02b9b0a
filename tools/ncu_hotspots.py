#!/usr/bin/env python
"""Per-instruction stall samples of a kernel from an .ncu-rep captured with --import-source on (source page, SASS view):
stall-reason shares of the kernel and the instructions that collect the most samples.

usage: python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep "(int)3,(int)3," > profiles/<name>.txt
(the second argument selects kernels whose name, spaces removed, contains it)
"""
import subprocess
import sys

rep = sys.argv[1]
open("/tmp/src_hot.csv", "w").write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                                                   capture_output=True, text=True).stdout)
sys.argv = [sys.argv[0]] + sys.argv[2:]
import csv, sys
# split per kernel
rows=list(csv.reader(open('/tmp/src_hot.csv')))
kern=None; data={}
hdr=None
for r in rows:
    if r and r[0]=="Kernel Name":
        kern=r[1]; data[kern]=[]; hdr=None; continue
    if r and r[0]=="Address": hdr=r; continue
    if kern and hdr and len(r)>=len(hdr)-2:
        data[kern].append(dict(zip(hdr,r)))
which=sys.argv[1] if len(sys.argv)>1 else "(int)3>" 
for k,v in data.items():
    if which not in k.replace(" ",""): 
        print("skip",k[:80], len(v)); continue
    print("==",k[:100], len(v))
    tot=sum(int(x["# Samples"]) for x in v)
    totinst=sum(int(x["Instructions Executed"]) for x in v)
    print("total samples",tot,"warp insts",totinst)
    stalls=[c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    agg={s:sum(int(x[s]) for x in v) for s in stalls}
    print({k2:round(v2/tot,3) for k2,v2 in sorted(agg.items(), key=lambda z:-z[1]) if v2>0})
    # top instructions by samples
    idx=sorted(range(len(v)), key=lambda i:-int(v[i]["# Samples"]))[:45]
    for i in sorted(idx):
        x=v[i]
        top=sorted(((int(x[s]),s) for s in stalls), reverse=True)[:2]
        print(i, x["Source"].strip()[:70].ljust(70), x["# Samples"], x["Instructions Executed"], top, x["L1 Wavefronts Shared"], x["L1 Wavefronts Shared Ideal"])
