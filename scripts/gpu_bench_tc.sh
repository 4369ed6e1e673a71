timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1f.json 2>gpurun_out/bench_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], r['achieved'], r['frac'], r['measured_gather_ceilings'], r['frac_of_l2_gather_64B'], r['traffic'])"
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
timeout 600 $B > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:shade_tc -c 2 -f -o gpurun_out/prof_shade_tc $B > gpurun_out/ncu_tc.log 2>&1; tail -2 gpurun_out/ncu_tc.log
