#!/usr/bin/env python
"""Two-kernel split probe (TVM_SPLIT_PROBE build): sigma march that emits the appearance-sample list + app_list_kernel.
    TVM_B200_LIB=iffnerf_b200/variants/libtvm_split_*.so python scripts/bench_split.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import _lib
from oracle import fixtures as fx
from tests import helpers as H
dev = torch.device("cuda:0")
fld = fx.make_field([300] * 3, density_shift=0.0)
m = H.module_from_field(fld, dev)
rays = fx.config2_rays().to(dev)
n, S = rays.shape[0], m.nSamples
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
probe = hasattr(raw, "tvm_split_probe_set")
m.mlp_precision = "fp32"
d, keep = m.field_desc()
need = C.c_size_t(0)
lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need))
ws = torch.zeros((need.value,), dtype=torch.uint8, device=dev)
bg = m._bg(None, True, dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
views = m.workspace_views(d, ws, n)
cap = 4 << 20
entries = torch.empty((cap * 8, 4), device=dev)
grp_ray = torch.empty((cap,), dtype=torch.int32, device=dev)
counter = torch.zeros((1,), dtype=torch.int32, device=dev)
if probe:
    raw.tvm_split_probe_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
    raw.tvm_split_probe_app.argtypes = [C.POINTER(_lib.FieldDesc), C.c_int64, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    assert raw.tvm_split_probe_set(entries.data_ptr(), grp_ray.data_ptr(), counter.data_ptr(), cap) == 0

def march():
    _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, None, _lib.ptr(bg), _lib.F_EARLY_TERM | _lib.F_NO_SHADE,
                                  None, None, None, None, None, None, None, None, None, _lib.ptr(ws), ws.numel(), st), "march")
def prep():
    counter.zero_(); views["ray_feat"].zero_()
def app(ctas):
    assert raw.tvm_split_probe_app(C.byref(d), n, ws.data_ptr(), ws.numel(), ctas, st) == 0

def timeit(fn, reps=6):
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / reps, 4)

out = {"lib": os.path.basename(_lib.LIB_PATH), "probe": probe}
if not probe:
    march(); march()
    out["march_ms"] = timeit(march)
    rf = views["ray_feat"]
    out["ray_feat_sum"] = float(rf.double().sum()); out["ray_feat_abs"] = float(rf.double().abs().sum())
else:
    prep(); march(); torch.cuda.synchronize()
    out["groups"] = int(counter.item())
    for ctas in (148 * 4, 148 * 8, 148 * 16):
        views["ray_feat"].zero_(); app(ctas); torch.cuda.synchronize()
        def only_app():
            app(ctas)
        out[f"app_ms_{ctas}"] = timeit(only_app)
    def sigma_only():
        counter.zero_(); march()
    out["sigma_ms"] = timeit(sigma_only)
    out["prep_ms"] = timeit(prep)
    def both():
        prep(); march(); app(148 * 8)
    out["split_total_ms"] = timeit(both)
    prep(); march(); app(148 * 8); torch.cuda.synchronize()
    rf = views["ray_feat"]
    out["ray_feat_sum"] = float(rf.double().sum()); out["ray_feat_abs"] = float(rf.double().abs().sum())
print(json.dumps(out))
