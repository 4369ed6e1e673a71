#!/usr/bin/env python
"""Hottest SASS instructions (stall samples) of the first launch whose kernel name contains a substring.
Usage: ncu_sass_hot.py rep kernel-substring [top]"""
import csv, io, subprocess, sys
def fl(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; take = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if hdr is not None: break
        take = want in r[1]; continue
    if not take: continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is not None and len(r) == len(hdr): data.append(r)
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(fl(r[iN]) for r in data)
hot = sorted(range(len(data)), key=lambda k: -fl(data[k][iN]))[:top]
for k in sorted(hot):
    print(f"{k:5d} samp={100*fl(data[k][iN])/tot:5.2f}% exec={fl(data[k][iI]):.2e}  {data[k][iS][:100]}")
