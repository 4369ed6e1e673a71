#!/bin/bash
# one ncu --set full capture of the backward kernels of one bench_bwd.py case: gpu_bwd_full.sh TAG [case]
mkdir -p gpurun_out
TAG=${1:-bwd}
CMD="python scripts/bench_bwd.py ${2:-train_65536}"
timeout 300 $CMD > gpurun_out/bench_bwd_${TAG}.log 2>&1 || exit 1
tail -1 gpurun_out/bench_bwd_${TAG}.log
ncu --set full --clock-control none --import-source on -k 'regex:app_bwd|march_bwd|march_fwd_kernel<1' -s 4 -c 4 -f -o gpurun_out/prof_bwd_${TAG} $CMD > gpurun_out/ncu_bwd_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_bwd_${TAG}.log
