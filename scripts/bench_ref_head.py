#!/usr/bin/env python
"""800x800 eval render with the `Ref` head (configs/lego.txt:25): fused tail kernel vs the torch-op tail."""
import contextlib, io, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import iffnerf_b200 as I
from iffnerf_b200 import synthetic as syn
dev = torch.device("cuda:0")
m = syn.build_model([300] * 3, dev, shading="Ref")
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
for name, flag in (("fused_tail_ms", True), ("torch_tail_ms", False)):
    m.ref_kernel = flag
    out[name] = timeit(lambda: m.render_eval(rays, white_bg=True))
m.native_shade_backup = None
print(json.dumps(out))
# train step with this head (4096 rays, S=1039): fused tail backward kernel vs torch autograd through the head
g = torch.Generator().manual_seed(0)
allrays = syn.config2_rays()
tr = allrays[torch.randint(0, allrays.shape[0], (4096,), generator=g)].to(dev)
target = torch.rand(4096, 3, device=dev)
ones = torch.ones(3, device=dev)
m.train()
m.ref_kernel = True
def train_step():
    m.zero_grad(set_to_none=True)
    rgb, _, _, alpha, _, _ = m(tr, bg_color=ones, is_train=True, N_samples=1039)
    (torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))).backward()
def timeit2(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / reps, 3)
res = {}
for name, flag in (("train_fused_ms", True), ("train_torch_tail_ms", False)):
    m.ref_kernel_train = flag
    res[name] = timeit2(train_step)
print(json.dumps(res))
