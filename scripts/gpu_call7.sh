#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sweep7.jsonl
for so in iffnerf_b200/variants/*.so; do
  TVM_B200_LIB=$PWD/$so timeout 300 python scripts/bench_march.py --steps 8 --march-only --tag $(basename $so .so) >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
done
timeout 300 python scripts/bench_march.py --steps 8 --march-only --tag default >> gpurun_out/sweep7.jsonl 2>> gpurun_out/sweep7.err
cat gpurun_out/sweep7.jsonl
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,sm__warps_active.avg.pct_of_peak_sustained_active
for v in noapp_b4 noapp_b6 noapp_b8; do
TVM_B200_LIB=$PWD/iffnerf_b200/variants/libtvm_$v.so ncu --metrics $M --clock-control none -k regex:march_fwd -s 2 -c 1 --csv --log-file gpurun_out/ncu7_$v.csv python scripts/bench_march.py --steps 1 --march-only > /dev/null 2>&1
echo $v; tail -7 gpurun_out/ncu7_$v.csv | cut -d, -f13-
done
