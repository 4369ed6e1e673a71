#!/usr/bin/env python
"""800x800 render with the reference's default head width (fea_pe = view_pe = 6, in_mlpC = 390): shading on the tensor
cores (K-chunked bf16x3 kernel) vs the fp32 SIMT kernel; march stage identical."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import synthetic as syn
dev = torch.device("cuda:0")
m = syn.build_model([300] * 3, dev, view_pe=6, fea_pe=6)
rays = syn.config2_rays().to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return round(tot / reps, 3)
out = {}
m.mlp_precision = "fp32"
ref = m.render_eval(rays, white_bg=True)["rgb_map"].clone()
out["step_ms_fp32_simt"] = timeit(lambda: m.render_eval(rays, white_bg=True))
m.mlp_precision = "auto"
out["auto_mode"] = m._shade_mode()
out["max_abs_rgb_tc3_vs_fp32"] = float((m.render_eval(rays, white_bg=True)["rgb_map"] - ref).abs().max())
out["step_ms_tc3"] = timeit(lambda: m.render_eval(rays, white_bg=True))
print(json.dumps(out))
