#!/usr/bin/env python
"""End-to-end (host rays -> host rgb/depth) timing of OctreeRender_trilinear_fast for different slice counts."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iffnerf_b200 as I
from oracle import fixtures as fx
from tests import helpers as H
dev = torch.device("cuda:0")
fld = fx.make_field([300] * 3, density_shift=0.0)
m = H.module_from_field(fld, dev)
rays = fx.config2_rays().pin_memory()
n = rays.shape[0]
out = torch.empty((n, 4)).pin_memory()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for prec in ("auto",):
    m.mlp_precision = prec
    for slices in (1, 2, 4, 6, 8, 16):
        m.host_ray_slices = slices
        def step():
            rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev)
            out[:, :3].copy_(rgb, non_blocking=True); out[:, 3].copy_(depth, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3): step()
        tot = 0.0; wall = 0.0
        for _ in range(10):
            flush.fill_(1); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record(); step(); e1.record(); torch.cuda.synchronize(); wall += time.perf_counter() - t0
            tot += e0.elapsed_time(e1)
        print(json.dumps({"mlp": prec, "slices": slices, "e2e_ms": round(tot / 10, 3), "wall_ms": round(wall * 100, 3)}))
