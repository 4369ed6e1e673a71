#!/usr/bin/env python
"""Basic-block view of an .ncu-rep SASS page: runs of instructions with equal execution count, sorted by total
warp-instructions.  Usage: ncu_sass_blocks.py rep [top] [kernel-substring]   (first launch whose name contains it)"""
import csv, io, subprocess, sys
def fl(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; take = False
for r in rows:
    if r and r[0] == "Kernel Name":
        if hdr is not None: break          # first matching launch only
        take = want in r[1]
        continue
    if not take: continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is not None and len(r) == len(hdr): data.append(r)
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
blocks = []; cur = None
for k, r in enumerate(data):
    c = fl(r[iI])
    if cur is None or c != cur["cnt"]:
        cur = {"start": k, "cnt": c, "n": 0, "samp": 0.0, "ops": {}}
        blocks.append(cur)
    cur["n"] += 1; cur["samp"] += fl(r[iN])
    op = r[iS].split()[0] if r[iS].split() else "?"
    if op.startswith("@"): op = r[iS].split()[1]
    op = op.split(".")[0]
    cur["ops"][op] = cur["ops"].get(op, 0) + 1
tot = sum(b["cnt"] * b["n"] for b in blocks); ts = sum(b["samp"] for b in blocks)
print(f"{len(data)} SASS instructions, {tot:.3e} warp-inst executed")
for b in sorted(blocks, key=lambda b: -b["cnt"] * b["n"])[:top]:
    ops = ", ".join(f"{k}:{v}" for k, v in sorted(b["ops"].items(), key=lambda kv: -kv[1])[:8])
    print(f"@{b['start']:5d} len={b['n']:4d} exec={b['cnt']:.3e} share={100*b['cnt']*b['n']/tot:5.1f}% samp={100*b['samp']/ts:5.1f}%  {ops}")
