#!/usr/bin/env python
"""Training-step timing on the bench field (BASELINE config 3: 4096-ray batch, S=1039, fwd+bwd, fp32).
Prints ms per phase (CUDA events) — a tuning aid; the parity of this step is tested in tests/test_gpu_grad.py."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--adam", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
m = syn.config2_model(dev)
m.train()
allrays = syn.config2_rays()
g = torch.Generator().manual_seed(0)
opt = torch.optim.Adam(m.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99)) if a.adam else None
ev = lambda: torch.cuda.Event(enable_timing=True)
tot = {"fwd": 0.0, "bwd": 0.0, "opt": 0.0}
for it in range(a.steps + 3):
    idx = torch.randint(0, allrays.shape[0], (a.rays,), generator=g)
    rays = allrays[idx].to(dev)
    target = torch.rand(a.rays, 3, device=dev)
    e = [ev() for _ in range(4)]
    m.zero_grad(set_to_none=True)
    e[0].record()
    rgb, depth, acc, alpha, z, dists = m(rays, bg_color=torch.ones(3, device=dev), is_train=True, N_samples=1039)
    loss = torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))
    e[1].record()
    loss.backward()
    e[2].record()
    if opt is not None:
        opt.step()
    e[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        tot["fwd"] += e[0].elapsed_time(e[1]); tot["bwd"] += e[1].elapsed_time(e[2]); tot["opt"] += e[2].elapsed_time(e[3])
out = {k: round(v / a.steps, 3) for k, v in tot.items()}
out["step_ms"] = round(sum(out.values()), 3)
out["rays"] = a.rays
out["rays_per_s"] = round(a.rays / (out["step_ms"] / 1e3))
print(json.dumps(out))
