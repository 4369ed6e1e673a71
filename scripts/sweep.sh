#!/bin/bash
# run scripts/bench_march.py over every variant library (GPU box)
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for so in iffnerf_b200/variants/*.so; do
  TVM_B200_LIB=$PWD/$so timeout 300 python scripts/bench_march.py --steps 5 --tag $(basename $so .so) >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
cat gpurun_out/sweep.jsonl
