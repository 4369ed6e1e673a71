#!/usr/bin/env python
"""End-to-end (host rays -> host rgb/depth) timing of OctreeRender_trilinear_fast for different slice counts, next to
the bare H2D / D2H copy times of the same buffers."""
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
rgb_h = torch.empty((n, 3)).pin_memory()
depth_h = torch.empty((n,)).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / reps, 3)

rd = torch.empty_like(rays, device=dev); rgb_d = torch.empty((n, 3), device=dev); dep_d = torch.empty((n,), device=dev)
print(json.dumps({"h2d_ms": timeit(lambda: rd.copy_(rays, non_blocking=True)),
                  "d2h_ms": timeit(lambda: (rgb_h.copy_(rgb_d, non_blocking=True), depth_h.copy_(dep_d, non_blocking=True))),
                  "device_ms": timeit(lambda: m.render_eval(rd, white_bg=True))}), flush=True)
for slices, frac in ((1, 1.0), (2, 1.0), (2, 0.5), (2, 0.25), (3, 0.5), (3, 0.3), (4, 0.5), (4, 0.3), (5, 0.5), (6, 0.5)):
    m.host_ray_slices = slices
    m.host_first_slice_frac = frac
    def step():
        rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev)
        rgb_h.copy_(rgb, non_blocking=True); depth_h.copy_(depth, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    def step_piped():           # downloads of a slice behind the kernels of the next (out_host)
        I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev, out_host=(rgb_h, depth_h))
        torch.cuda.current_stream().synchronize()
    print(json.dumps({"slices": slices, "first_frac": frac, "e2e_ms": timeit(step), "e2e_out_host_ms": timeit(step_piped)}),
          flush=True)
