#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/split.jsonl
timeout 200 python scripts/bench_split.py >> gpurun_out/split.jsonl 2>> gpurun_out/split.err
for so in iffnerf_b200/variants/libtvm_split_*.so; do
  TVM_B200_LIB=$PWD/$so timeout 200 python scripts/bench_split.py >> gpurun_out/split.jsonl 2>> gpurun_out/split.err
done
cat gpurun_out/split.jsonl; tail -5 gpurun_out/split.err
