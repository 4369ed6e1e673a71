#!/bin/bash
# full round check (tests, smoke, bench, ncu launch list + full capture) + a full capture of the train-step kernels
bash scripts/gpu_check.sh r1j
timeout 300 python scripts/bench_train.py > gpurun_out/train_r1j.json 2> gpurun_out/train_r1j.err; cat gpurun_out/train_r1j.json
ncu --set full --clock-control none --import-source on -k 'regex:march_bwd|shade_bwd|march_fwd' -s 12 -c 4 -f -o gpurun_out/prof_train_r1j python scripts/bench_train.py --steps 3 > gpurun_out/ncu_train_r1j.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
