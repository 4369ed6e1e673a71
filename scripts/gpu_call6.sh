#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  (cd _old && timeout 300 python scripts/bench_train.py) > gpurun_out/train_old_$i.json 2> gpurun_out/train_old_$i.err; cat gpurun_out/train_old_$i.json
  timeout 300 python scripts/bench_train.py > gpurun_out/train_new_$i.json 2> gpurun_out/train_new_$i.err; cat gpurun_out/train_new_$i.json
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_grad.py -m gpu -x -q 2>&1 | tail -2
