#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c3.log
tail -15 gpurun_out/pytest_c3.log
timeout 600 python scripts/bench_e2e.py > gpurun_out/e2e_sweep.jsonl 2> gpurun_out/e2e_sweep.err; cat gpurun_out/e2e_sweep.jsonl; tail -3 gpurun_out/e2e_sweep.err
timeout 300 python scripts/bench_l1_gather.py > gpurun_out/l1_gather.jsonl 2> gpurun_out/l1_gather.err; cat gpurun_out/l1_gather.jsonl; tail -3 gpurun_out/l1_gather.err
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum
ncu --metrics $M --clock-control none -k regex:gather_bench -c 40 --csv --log-file gpurun_out/ncu_l1_gather.csv python scripts/bench_l1_gather.py > /dev/null 2>&1
