#!/usr/bin/env python
"""Where does the end-to-end overhead go? H2D / kernels / D2H timed separately on the bench workload."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iffnerf_b200 as I
from iffnerf_b200 import synthetic as syn
dev = torch.device("cuda:0")
m = syn.config2_model(dev)
rays = syn.config2_rays().pin_memory()
n = rays.shape[0]
out = torch.empty((n, 4)).pin_memory()
rgb = torch.empty((n, 3), device=dev); depth = torch.empty(n, device=dev)
def ev(): return torch.cuda.Event(enable_timing=True)
res = {}
for name, fn in (("h2d", lambda: rays.to(dev, non_blocking=True)),
                 ("d2h", lambda: (out[:, :3].copy_(rgb, non_blocking=True), out[:, 3].copy_(depth, non_blocking=True))),
                 ("render_dev", lambda: m.render_eval(rays_d, white_bg=True)),
                 ("octree_host", lambda: I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev))):
    rays_d = rays.to(dev)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev(); t0 = time.perf_counter(); e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize(); res[name + "_ms"] = round(e0.elapsed_time(e1) / 10, 3); res[name + "_wall_ms"] = round((time.perf_counter() - t0) * 100, 3)
# contiguous D2H variant: one [n,4] device tensor -> one copy
packed = torch.empty((n, 4), device=dev)
def d2h_one(): out.copy_(packed, non_blocking=True)
for _ in range(3): d2h_one()
torch.cuda.synchronize(); e0, e1 = ev(), ev(); e0.record()
for _ in range(10): d2h_one()
e1.record(); torch.cuda.synchronize(); res["d2h_contiguous_ms"] = round(e0.elapsed_time(e1) / 10, 3)
print(json.dumps(res))
