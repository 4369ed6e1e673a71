#!/usr/bin/env python
"""Device-side timing of tvm_march_bwd alone (tuning aid): train shape (4096 random rays, S=1039, factor-gradient
scatter) and pose shape (64x1024 rays, S=1036, d(rays) only).   TVM_B200_LIB=variant.so python scripts/bench_bwd.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import _lib
from iffnerf_b200 import synthetic as syn
dev = torch.device("cuda:0")
m = syn.config2_model(dev)
lib = _lib.load()
allrays = syn.config2_rays()
g = torch.Generator().manual_seed(0)
d, keep = m.field_desc()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
bg = m._bg(None, True, dev)
out = {"lib": os.path.basename(_lib.LIB_PATH)}

ONLY = sys.argv[1:]          # optional: tags to run (default all)


def run(n, S, scatter, pose, tag, flags=0):
    if ONLY and tag not in ONLY:
        return
    rays = allrays[torch.randint(0, allrays.shape[0], (n,), generator=g)].to(dev)
    need = C.c_size_t(0)
    lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
    _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, None, _lib.ptr(bg), _lib.F_NO_SHADE,
                                  None, None, None, None, None, None, None, None, None, _lib.ptr(ws), ws.numel(), st), "fwd")
    ta = sum(m.app_n_comp)
    d_feat = torch.randn((n, ta), device=dev) * 1e-3
    d_acc = torch.randn((n,), device=dev) * 1e-3
    d_alpha = (torch.randn((n, S), device=dev) * 1e-5) if scatter else None
    g_fac = torch.zeros(int(d.n_factor_floats), device=dev) if scatter else None
    g_rays = torch.zeros((n, 6), device=dev) if pose else None
    def bwd():
        _lib.check(lib.tvm_march_bwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, None, flags, _lib.ptr(d_feat), _lib.ptr(d_acc),
                                     _lib.ptr(d_alpha), _lib.ptr(g_fac), _lib.ptr(g_rays), _lib.ptr(ws), ws.numel(), st), "bwd")
    for _ in range(3): bwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): bwd()
    e1.record(); torch.cuda.synchronize()
    out[tag + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)

run(4096, 1039, True, False, "train_4096")
run(65536, 1036, False, True, "pose_65536")
run(65536, 1039, True, False, "train_65536")
run(4096, 1039, True, False, "train_4096_runs", _lib.F_BWD_RUNS)
run(65536, 1039, True, False, "train_65536_runs", _lib.F_BWD_RUNS)
print(json.dumps(out))
