#!/bin/bash
# bench.py at N ranks under torchrun (as the driver launches it); writes gpurun_out/bench_n${N}.json
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "exit $?"; tail -2 gpurun_out/bench_n${N}.err | cut -c1-200
python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['bf16_mlp_mode']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-160
