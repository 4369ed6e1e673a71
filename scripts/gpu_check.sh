#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + one full capture of the march kernels.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_${TAG}.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; tail -2 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench exit $?"
cat gpurun_out/bench_${TAG}.json | cut -c1-3000
tail -5 gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "reference arm exit $?"
cut -c1-400 gpurun_out/bench_ref_${TAG}.json
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 $BENCH > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 600 $BENCH > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:march_fwd|app_gather|shade_tc3' -s 4 -c 4 -f -o gpurun_out/prof_march_${TAG} $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
