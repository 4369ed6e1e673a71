#!/usr/bin/env python
"""Quick device-side timing of the march / shade kernels on the bench workload (tuning aid, not the official bench).
    TVM_B200_LIB=path/to/variant.so python scripts/bench_march.py [--steps K] [--no-early]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iffnerf_b200 import _lib
from iffnerf_b200 import synthetic as syn

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-early", action="store_true")
ap.add_argument("--tag", default=os.environ.get("TVM_B200_LIB", "default"))
ap.add_argument("--tile", default="", help="WxH: reorder the 800x800 rays so 32 consecutive rays form a WxH pixel tile (locality probe)")
ap.add_argument("--march-only", action="store_true")
ap.add_argument("--fused", action="store_true", help="one-kernel march (no TVM_F_SPLIT_APP)")
a = ap.parse_args()
dev = torch.device("cuda:0")
m = syn.config2_model(dev)
rays = syn.config2_rays()
if a.tile:
    tw, th = (int(v) for v in a.tile.split("x"))
    img = rays.view(800, 800, -1)
    rays = img.view(800 // th, th, 800 // tw, tw, -1).permute(0, 2, 1, 3, 4).reshape(-1, img.shape[-1]).contiguous()
rays = rays.to(dev)
n, S = rays.shape[0], m.nSamples
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
m.mlp_precision = "fp32"
d, keep = m.field_desc()
need = C.c_size_t(0)
lib.tvm_workspace_bytes(C.byref(d), n, 0 if a.fused else _lib.F_SPLIT_APP, C.byref(need))
ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
bg = m._bg(None, True, dev)
rgb = torch.empty((n, 3), device=dev); depth = torch.empty(n, device=dev); acc = torch.empty(n, device=dev)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
fl = (0 if a.no_early else _lib.F_EARLY_TERM) | (0 if a.fused else (_lib.F_SPLIT_APP | _lib.F_ZERO_UNLIT))

def march():
    _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, None, _lib.ptr(bg), fl | _lib.F_NO_SHADE,
                                  None, None, None, None, None, None, None, None, None, _lib.ptr(ws), ws.numel(), st), "march")
def shade():
    _lib.check(lib.tvm_shade_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], _lib.ptr(bg), 0, _lib.ptr(rgb), _lib.ptr(depth),
                                 _lib.ptr(acc), _lib.ptr(ws), ws.numel(), st), "shade")
def timeit(fn):
    fn(); fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(a.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / a.steps
m.mlp_precision = "bf16"
d2, keep2 = m.field_desc()
def shade_tc():
    _lib.check(lib.tvm_shade_fwd(C.byref(d2), _lib.ptr(rays), n, rays.shape[1], _lib.ptr(bg), _lib.F_MLP_BF16, _lib.ptr(rgb),
                                 _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(ws), ws.numel(), st), "shade_tc")
if a.march_only:
    out = {"tag": a.tag, "tile": a.tile, "fused": a.fused, "march_ms": round(timeit(march), 4)}
    v = m.workspace_views(d, ws, n)      # checksums of the march outputs (compare builds)
    out.update(feat_sum=float(v["ray_feat"].double().sum()), feat_abs=float(v["ray_feat"].double().abs().sum()),
               acc_sum=float(v["acc"].double().sum()), depth_sum=float(v["depth"].double().sum()),
               app=int(v["app_count"].sum()), sigma=int(v["sigma_count"].sum()))
    print(json.dumps(out))
    sys.exit(0)
march()
ref_rgb = torch.empty_like(rgb); shade(); ref_rgb.copy_(rgb)
tc_ms = round(timeit(shade_tc), 4)
tc_err = float((rgb - ref_rgb).abs().max())
m.mlp_precision = "tc3"
d3, keep3 = m.field_desc()
def shade_tc3():
    _lib.check(lib.tvm_shade_fwd(C.byref(d3), _lib.ptr(rays), n, rays.shape[1], _lib.ptr(bg), _lib.F_MLP_TC3, _lib.ptr(rgb),
                                 _lib.ptr(depth), _lib.ptr(acc), _lib.ptr(ws), ws.numel(), st), "shade_tc3")
tc3_ms = round(timeit(shade_tc3), 4)
tc3_err = float((rgb - ref_rgb).abs().max())
print(json.dumps({"shade_tc3_ms": tc3_ms, "shade_tc3_max_abs_vs_fp32": tc3_err}))
print(json.dumps({"tag": a.tag, "shade_tc_ms": tc_ms, "shade_tc_max_abs_vs_fp32": tc_err, "march_ms": round(timeit(march), 4), "shade_ms": round(timeit(shade), 4), "early": not a.no_early}))
