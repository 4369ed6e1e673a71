#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graphs.py tests/test_gpu_grad.py tests/test_gpu_raygen.py -m gpu -x -q 2>&1 | tail -25
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c10.json 2> gpurun_out/bench_c10.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c10.json')); print(d['value'], d['e2e']['value'], json.dumps(d['other_configs'], indent=1))"; tail -5 gpurun_out/bench_c10.err
