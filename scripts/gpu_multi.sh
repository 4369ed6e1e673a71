#!/bin/bash
# multi-GPU check (run under `gpurun --gpus N`): NCCL / peer-memory parity test + the sharded bench at N GPUs
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_${TAG}_n${N}.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_${TAG}.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_multi_${TAG}.log
tail -15 gpurun_out/pytest_multi_${TAG}.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n${N}.json 2> gpurun_out/bench_${TAG}_n${N}.err; echo "bench exit $?"
cut -c1-2500 gpurun_out/bench_${TAG}_n${N}.json
tail -15 gpurun_out/bench_${TAG}_n${N}.err
