#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/bwd_sweep.jsonl
for so in iffnerf_b200/variants/libtvm_bwd_*.so; do
  TVM_B200_LIB=$PWD/$so timeout 200 python scripts/bench_bwd.py >> gpurun_out/bwd_sweep.jsonl 2>> gpurun_out/bwd_sweep.err
done
cat gpurun_out/bwd_sweep.jsonl; tail -3 gpurun_out/bwd_sweep.err
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c9.json 2> gpurun_out/bench_c9.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c9.json')); print(d['value'], d['e2e']['value'], json.dumps(d['other_configs']))"; tail -3 gpurun_out/bench_c9.err
