#!/bin/bash
# ncu of the backward kernels at 65 536 rays (scripts/bench_bwd.py): launch list + one full capture per kernel
mkdir -p gpurun_out
TAG=${1:-bwd}
timeout 300 python scripts/bench_bwd.py > gpurun_out/bench_bwd_${TAG}.log 2>&1 || exit 1
tail -1 gpurun_out/bench_bwd_${TAG}.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bwd_${TAG}.csv python scripts/bench_bwd.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_bwd_${TAG}.csv", errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ki, vi, gi = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value"), rows[hdr].index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    k = (r[ki][:90], r[gi])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
for (k, g), (n, t) in agg.items():
    print(f"{n:4d} x {t / n / 1e3:9.1f} us  grid={g:>16s}  {k}")
PY
