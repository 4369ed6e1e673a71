#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo). Usage: ncu_source.py rep [top]"""
import csv, io, subprocess, sys

def fl(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs, cur, path = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        path = r[1]
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": [], "path": path}
        secs.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
# merge the same (file,line) over sections belonging to the first launch only
agg = {}
seen_paths = set()
for s in secs:
    if s["path"] in seen_paths:
        continue
    seen_paths.add(s["path"])
    h = s["hdr"]
    iL, iI, iS = h.index("Line No"), h.index("Instructions Executed"), h.index("# Samples")
    for r in s["rows"]:
        key = (s["path"].split("/")[-1], r[iL])
        a = agg.setdefault(key, [0.0, 0.0, r[1]])
        a[0] += fl(r[iI]); a[1] += fl(r[iS])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"total warp-inst {ti:.3e}, stall samples {ts:.0f}")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:18s}:{l:>4s} inst={100*a[0]/ti:5.1f}% samp={100*a[1]/ts:5.1f}%  {a[2].strip()[:100]}")
