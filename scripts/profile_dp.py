#!/usr/bin/env python
"""Data-parallel train step (config 3, 4096 rays per rank) under torch.profiler: per-step device time of every kernel on
rank 0.  Run under torchrun:  python -m torch.distributed.run --nproc-per-node N scripts/profile_dp.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from iffnerf_b200 import sharding, synthetic as syn

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
m = syn.config2_model(dev)
m.train()
allrays = syn.config2_rays()
g = torch.Generator().manual_seed(100 + rank)
rays = allrays[torch.randint(0, allrays.shape[0], (4096,), generator=g)].to(dev)
target = torch.rand(4096, 3, device=dev)
jit = torch.rand(4096, device=dev)
ones = torch.ones(3, device=dev)
sync = sharding.GradSync(m, average=True).install() if world > 1 else None

def step():
    m.zero_grad(set_to_none=True)
    rgb, _, _, alpha, _, _ = m(rays, bg_color=ones, is_train=True, N_samples=1039, jitter=jit)
    (torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))).backward()
    if sync is not None:
        sync.finish()
for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
N = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
if rank == 0:
    rows = [(e.key[:70], e.device_time_total / N, e.count / N) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[1])
    print(json.dumps({"world": world, "ms_per_step": ms, "kernels_us_per_step": [[k, round(t, 1), c] for k, t, c in rows[:18]],
                      "sum_us": round(sum(r[1] for r in rows), 1)}))
if world > 1:
    dist.destroy_process_group()
