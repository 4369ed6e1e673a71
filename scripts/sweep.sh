#!/bin/bash
# run scripts/bench_march.py over the default library and every variant library; per-kernel times via an ncu launch list
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for so in iffnerf_b200/libtvm_b200.so iffnerf_b200/variants/*.so; do
  TVM_B200_LIB=$PWD/$so timeout 300 python scripts/bench_march.py --march-only --steps 10 --tag $(basename $so .so) >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
  TVM_B200_LIB=$PWD/$so ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:march_fwd|app_gather' --csv --log-file gpurun_out/times_$(basename $so .so).csv python scripts/bench_march.py --march-only --steps 3 > /dev/null 2>&1
done
cat gpurun_out/sweep.jsonl
