#!/bin/bash
# full GPU suite on the split march; pixel-tile locality probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_suite.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_suite.log
tail -8 gpurun_out/pytest_suite.log
: > gpurun_out/march_suite.jsonl
for t in "" "--tile 8x4" "--tile 4x8" "--tile 16x2" "--tile 2x16"; do
  timeout 300 python scripts/bench_march.py --march-only --steps 10 $t >> gpurun_out/march_suite.jsonl 2>> gpurun_out/march_suite.err
done
cat gpurun_out/march_suite.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:march_fwd|app_gather' --csv --log-file gpurun_out/times_suite_tile8x4.csv python scripts/bench_march.py --march-only --steps 3 --tile 8x4 > /dev/null 2>&1
tail -5 gpurun_out/march_suite.err
