#!/usr/bin/env python
"""Pose-mode step (BASELINE config 5: 64 x 1024 rays, frozen field, gradient w.r.t. the rays) — run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel split, or plainly for the step time."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import synthetic as syn
dev = torch.device("cuda:0")
m = syn.config2_model(dev)
m.eval()
for p in m.parameters():
    p.requires_grad_(False)
m.eval_sample_outputs = "--samples" in sys.argv
g = torch.Generator().manual_seed(0)
allrays = syn.config2_rays()
pick = torch.randint(0, allrays.shape[0], (64 * 1024,), generator=g)
if "--sorted" in sys.argv:          # locality probe: the same rays in pixel order
    pick = pick.sort().values
if "--tiled" in sys.argv:           # ... in 8x4-pixel tile order
    y, x = pick // 800, pick % 800
    pick = pick[torch.argsort(((y // 4) * 100 + x // 8) * 32 + (y % 4) * 8 + x % 8)]
prays = allrays[pick].to(dev)
target = torch.rand(64 * 1024, 3, device=dev)
bg = torch.rand(3, device=dev)
def step():
    r = prays.clone().requires_grad_(True)
    rgb = m(r, bg_color=bg, is_train=False)[0]
    torch.mean((rgb - target) ** 2).backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"order": "sorted" if "--sorted" in sys.argv else ("tiled" if "--tiled" in sys.argv else "random"), "sample_outputs": m.eval_sample_outputs, "pose_step_ms": round(e0.elapsed_time(e1) / 5, 3)}))
