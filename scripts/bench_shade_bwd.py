#!/usr/bin/env python
"""Device-side timing of tvm_shade_bwd alone (tuning aid): pose mode (d_view, no parameter gradients) and train mode
(basis + MLP gradients) at 1024 / 4096 / 65536 rays.   TVM_B200_LIB=variant.so python scripts/bench_shade_bwd.py"""
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

def run(n, train, tag):
    rays = allrays[torch.randint(0, allrays.shape[0], (n,), generator=g)].to(dev)
    need = C.c_size_t(0)
    lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need))
    ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
    _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], 1036, None, _lib.ptr(bg), _lib.F_NO_SHADE,
                                  None, None, None, None, None, None, None, None, None, _lib.ptr(ws), ws.numel(), st), "fwd")
    ta = sum(m.app_n_comp)
    g_rgb = torch.randn((n, 3), device=dev)
    d_feat = torch.empty((n, ta), device=dev)
    d_acc = torch.empty((n,), device=dev)
    d_view = None if train else torch.empty((n, 3), device=dev)
    g_basis = torch.zeros_like(m.basis_mat.weight) if train else None
    g_mlp = torch.zeros(int(lib.tvm_mlp_grad_floats(C.byref(d))), device=dev) if train else None
    def bwd():
        _lib.check(lib.tvm_shade_bwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], _lib.ptr(bg), _lib.ptr(g_rgb), None,
                                     _lib.ptr(d_feat), _lib.ptr(d_acc), _lib.ptr(g_basis), _lib.ptr(g_mlp), _lib.ptr(d_view),
                                     _lib.ptr(ws), ws.numel(), st), "shade_bwd")
    for _ in range(3): bwd()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): bwd()
    e1.record(); torch.cuda.synchronize()
    out[tag + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)

for n in (1024, 4096, 65536):
    run(n, False, f"pose_{n}")
    run(n, True, f"train_{n}")
print(json.dumps(out))
