#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grad.py tests/test_gpu_canary.py tests/test_gpu_train_loop.py tests/test_gpu_graphs.py tests/test_gpu_raygen.py -x -q > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_bwd.log
tail -6 gpurun_out/pytest_bwd.log
for n in 4096 65536; do timeout 300 python scripts/bench_train.py --rays $n --steps 10; done > gpurun_out/train_bwd.jsonl 2>&1
cat gpurun_out/train_bwd.jsonl
timeout 300 python scripts/bench_bwd.py > gpurun_out/bench_bwd.log 2>&1; tail -8 gpurun_out/bench_bwd.log
