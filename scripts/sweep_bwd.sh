#!/bin/bash
# scripts/bench_bwd.py over the default library and every variant library
mkdir -p gpurun_out
: > gpurun_out/sweep_bwd.jsonl
for so in iffnerf_b200/libtvm_b200.so iffnerf_b200/variants/*.so; do
  TVM_B200_LIB=$PWD/$so timeout 300 python scripts/bench_bwd.py "$@" 2>> gpurun_out/sweep_bwd.err | tail -1 >> gpurun_out/sweep_bwd.jsonl
done
cat gpurun_out/sweep_bwd.jsonl
