#!/bin/bash
# gpurun call: GPU parity tests on the default library, then march-kernel timing over variants / tile orders
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s2.log
tail -4 gpurun_out/pytest_s2.log
: > gpurun_out/sweep2.jsonl
for so in iffnerf_b200/variants/*.so; do
  TVM_B200_LIB=$PWD/$so timeout 300 python scripts/bench_march.py --steps 8 --march-only --tag $(basename $so .so) >> gpurun_out/sweep2.jsonl 2>> gpurun_out/sweep2.err
done
for t in 8x4 4x8 16x2 2x16; do
  timeout 300 python scripts/bench_march.py --steps 8 --march-only --tile $t --tag default >> gpurun_out/sweep2.jsonl 2>> gpurun_out/sweep2.err
done
cat gpurun_out/sweep2.jsonl
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread
ncu --metrics $M --clock-control none -k regex:march_fwd -s 2 -c 1 --csv --log-file gpurun_out/ncu_s2_new.csv python scripts/bench_march.py --steps 1 --march-only > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:march_fwd -s 2 -c 1 --csv --log-file gpurun_out/ncu_s2_tile.csv python scripts/bench_march.py --steps 1 --march-only --tile 8x4 > /dev/null 2>&1
tail -3 gpurun_out/ncu_s2_new.csv gpurun_out/ncu_s2_tile.csv | cut -c1-400
