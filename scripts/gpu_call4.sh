#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_c4.log
tail -12 gpurun_out/pytest_c4.log
timeout 300 python scripts/bench_march.py --steps 8 --march-only --tag pitch > gpurun_out/march_c4.jsonl 2> gpurun_out/march_c4.err; cat gpurun_out/march_c4.jsonl; tail -2 gpurun_out/march_c4.err
timeout 300 python scripts/bench_ref_head.py > gpurun_out/ref_head_c4.json 2> gpurun_out/ref_head_c4.err; cat gpurun_out/ref_head_c4.json; tail -2 gpurun_out/ref_head_c4.err
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,l1tex__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
ncu --metrics $M --clock-control none -k regex:march_fwd -s 2 -c 1 --csv --log-file gpurun_out/ncu_c4_march.csv python scripts/bench_march.py --steps 1 --march-only > /dev/null 2>&1
tail -9 gpurun_out/ncu_c4_march.csv | cut -d, -f13-
