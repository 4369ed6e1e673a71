#!/bin/bash
# one ncu --set full capture of the march kernels of scripts/bench_march.py (tuning aid): gpu_ncu_march.sh TAG [lib.so]
mkdir -p gpurun_out
TAG=${1:-m}
[ -n "$2" ] && export TVM_B200_LIB=$PWD/$2
CMD="python scripts/bench_march.py --march-only --steps 2"
timeout 300 $CMD > gpurun_out/march_${TAG}.log 2>&1 || { tail -5 gpurun_out/march_${TAG}.log; exit 1; }
tail -1 gpurun_out/march_${TAG}.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k 'regex:march_fwd' -s 4 -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
