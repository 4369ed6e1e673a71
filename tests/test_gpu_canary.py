"""Out-of-bounds guard for the kernels' outputs (compute-sanitizer is closed on this pool): every output is a slice in
the middle of a sentinel-filled buffer; after the call the sentinels on both sides must be untouched.  Ragged sizes
around the 32-ray march CTA, the 64-ray SIMT tile and the 128-ray tensor-core tile, lit and all-background batches."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from tests import helpers as H

pytestmark = pytest.mark.gpu
SENT = -12345.0
PAD = 1024


class Guarded:
    def __init__(self, shape, dev, dtype=torch.float32):
        n = int(np.prod(shape))
        self.buf = torch.full((n + 2 * PAD,), SENT if dtype.is_floating_point else 7, dtype=dtype, device=dev)
        self.view = self.buf[PAD:PAD + n].view(*shape)

    def intact(self):
        return bool((self.buf[:PAD] == SENT).all() and (self.buf[-PAD:] == SENT).all())


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


SIZES = (1, 31, 33, 63, 65, 127, 129, 4097)


@pytest.mark.parametrize("precision", ["auto", "fp32", "bf16"])
def test_render_eval_outputs_stay_inside_their_slices(dev, precision):
    fld, rays = fx.config1(0.0, "sphere", 7)
    m = H.module_from_field(fld, dev)
    m.mlp_precision = precision
    lit = rays[4000:4000 + max(SIZES)].to(dev)                  # centre of the image: hits the object
    away = lit.clone()
    away[:, 3:6] = -away[:, 3:6]                                   # looking away: background only (tile-skip path)
    for batch in (lit, away):
        for n in SIZES:
            rgb, depth = Guarded((n, 3), dev), Guarded((n,), dev)
            o = m.render_eval(batch[:n], white_bg=True, out_rgb=rgb.view, out_depth=depth.view)
            torch.cuda.synchronize()
            assert rgb.intact() and depth.intact(), (precision, n)
            assert torch.isfinite(rgb.view).all() and torch.isfinite(depth.view).all() and torch.isfinite(o["acc_map"]).all()
            assert not (rgb.view == SENT).any() and not (depth.view == SENT).any()      # every element was written
    assert float(m.render_eval(away[:256], white_bg=True)["rgb_map"].min()) == 1.0       # pure background is white


def test_point_queries_and_ray_generation_stay_inside_their_buffers(dev, built_lib):
    from iffnerf_b200 import _lib
    fld, rays = fx.config1(0.0, "sphere", 6)
    m = H.module_from_field(fld, dev)
    d, keep = m.field_desc()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    g = torch.Generator().manual_seed(0)
    kinv = (C.c_float * 9)(0.01, 0.0, -0.5, 0.0, 0.01, -0.5, 0.0, 0.0, 1.0)
    c2w = torch.eye(4, device=dev)
    for n in SIZES:
        pts = (torch.rand(n, 3, generator=g) * 2 - 1).to(dev)
        feat, dens = Guarded((n, 27), dev), Guarded((n,), dev)
        _lib.check(built_lib.tvm_point_appfeature(C.byref(d), _lib.ptr(pts), n, _lib.ptr(feat.view), st), "appfeature")
        _lib.check(built_lib.tvm_point_density(C.byref(d), _lib.ptr(pts), n, 0, 1.0, _lib.ptr(dens.view), st), "density")
        pix = torch.randint(0, 100, (n, 2), generator=g).to(device=dev, dtype=torch.int32)
        out = Guarded((n, 7), dev)
        _lib.check(built_lib.tvm_pixel_rays_fwd(_lib.ptr(c2w), 16, kinv, _lib.ptr(pix), None, 0, n, 3, _lib.ptr(out.view), st),
                   "pixel_rays")
        gc2w = Guarded((1, 3, 4), dev)
        gc2w.view.zero_()
        up = torch.randn(n, 7, generator=g).to(dev)
        _lib.check(built_lib.tvm_pixel_rays_bwd(_lib.ptr(c2w), 16, kinv, _lib.ptr(pix), None, 0, n, 3, _lib.ptr(up), 7,
                                                _lib.ptr(gc2w.view), st), "pixel_rays_bwd")
        torch.cuda.synchronize()
        assert feat.intact() and dens.intact() and out.intact() and gc2w.intact(), n
        for t in (feat.view, dens.view, out.view, gc2w.view):
            assert torch.isfinite(t).all() and not (t == SENT).any()


def test_ref_head_tail_kernel_stays_inside_its_slices(dev):
    from tests.test_gpu_ref_head import _ref_model
    m = _ref_model(dev)
    _, rays = fx.config1(0.0, None, 7)
    lit = rays[4000:4000 + max(SIZES)].to(dev)
    for n in SIZES:
        rgb, depth = Guarded((n, 3), dev), Guarded((n,), dev)
        m.render_eval(lit[:n], white_bg=True, out_rgb=rgb.view, out_depth=depth.view)
        torch.cuda.synchronize()
        assert rgb.intact() and depth.intact(), n
        assert torch.isfinite(rgb.view).all() and not (rgb.view == SENT).any()


def test_backward_kernels_stay_inside_their_buffers(dev, built_lib):
    """tvm_shade_bwd + tvm_march_bwd (factor-gradient scatter and d(rays)) with every output guarded."""
    from iffnerf_b200 import _lib
    fld, rays = fx.config1(0.0, "sphere", 6)
    m = H.module_from_field(fld, dev)
    m.mlp_precision = "fp32"
    d, keep = m.field_desc()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    S = m.nSamples
    ta = sum(m.app_n_comp)
    bg = torch.ones(3, device=dev)
    g = torch.Generator().manual_seed(1)
    for n in (1, 33, 65, 129, 1025):
        r = rays[4545:4545 + n].to(dev).contiguous()                 # starts at the image centre: hits the object
        need = C.c_size_t(0)
        built_lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need))
        ws = Guarded((need.value,), dev, dtype=torch.uint8)
        rgb, depth, acc = torch.empty(n, 3, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev)
        alpha = Guarded((n, S), dev)
        _lib.check(built_lib.tvm_render_fwd(C.byref(d), _lib.ptr(r), n, 6, S, None, _lib.ptr(bg), 0, _lib.ptr(rgb), _lib.ptr(depth),
                                            _lib.ptr(acc), _lib.ptr(alpha.view), None, None, None, None, None,
                                            _lib.ptr(ws.view), need.value, st), "fwd")
        d_rgb = torch.randn(n, 3, generator=g).to(dev)
        d_feat, d_acc, d_view = Guarded((n, ta), dev), Guarded((n,), dev), Guarded((n, 3), dev)
        g_basis = Guarded((27, ta), dev)
        g_mlp = Guarded((int(built_lib.tvm_mlp_grad_floats(C.byref(d))),), dev)
        g_basis.view.zero_(); g_mlp.view.zero_()
        _lib.check(built_lib.tvm_shade_bwd(C.byref(d), _lib.ptr(r), n, 6, _lib.ptr(bg), _lib.ptr(d_rgb), None, _lib.ptr(d_feat.view),
                                           _lib.ptr(d_acc.view), _lib.ptr(g_basis.view), _lib.ptr(g_mlp.view), _lib.ptr(d_view.view),
                                           _lib.ptr(ws.view), need.value, st), "shade_bwd")
        g_fac = Guarded((int(d.n_factor_floats),), dev)
        g_fac.view.zero_()
        g_rays = Guarded((n, 6), dev)
        g_rays.view.zero_()
        d_alpha = (torch.randn(n, S, generator=g) * 1e-3).to(dev)
        _lib.check(built_lib.tvm_march_bwd(C.byref(d), _lib.ptr(r), n, 6, S, None, 0, _lib.ptr(d_feat.view), _lib.ptr(d_acc.view),
                                           _lib.ptr(d_alpha), _lib.ptr(g_fac.view), _lib.ptr(g_rays.view), _lib.ptr(ws.view),
                                           need.value, st), "march_bwd")
        torch.cuda.synchronize()
        assert (ws.buf[:PAD] == 7).all() and (ws.buf[-PAD:] == 7).all(), n
        for name, t in (("alpha", alpha), ("d_feat", d_feat), ("d_acc", d_acc), ("d_view", d_view), ("g_basis", g_basis),
                        ("g_mlp", g_mlp), ("g_fac", g_fac), ("g_rays", g_rays)):
            assert t.intact(), (name, n)
            assert torch.isfinite(t.view).all(), (name, n)
        assert float(g_fac.view.abs().sum()) > 0 and float(g_rays.view.abs().sum()) > 0


def test_ref_head_backward_kernel_stays_inside_its_buffers(dev, built_lib):
    """tvm_shade_ref_bwd with every output guarded, ragged ray counts around its 64-ray tile."""
    from iffnerf_b200 import _lib
    from tests.test_gpu_ref_head import _ref_model
    m = _ref_model(dev)
    d, keep = m.field_desc()
    h, buf = m.packed_ref_head()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _, rays = fx.config1(0.0, None, 7)
    ta = sum(m.app_n_comp)
    bg = torch.ones(3, device=dev)
    g = torch.Generator().manual_seed(2)
    for n in (1, 63, 65, 129, 1000):
        r = rays[4545:4545 + n].to(dev).contiguous()
        o = m.render_eval(r, white_bg=True, keep_workspace=True, want_counts=True)
        v = o["workspace"]
        d_rgb = torch.randn(n, 3, generator=g).to(dev)
        d_feat, d_acc, d_view = Guarded((n, ta), dev), Guarded((n,), dev), Guarded((n, 3), dev)
        g_basis, g_par = Guarded((27, ta), dev), Guarded((buf.numel(),), dev)
        g_basis.view.zero_(); g_par.view.zero_()
        _lib.check(built_lib.tvm_shade_ref_bwd(C.byref(d), C.byref(h), _lib.ptr(r), n, 7, _lib.ptr(bg), _lib.ptr(v["ray_feat"]),
                                               _lib.ptr(v["acc"]), _lib.ptr(v["app_count"]), _lib.ptr(d_rgb), None,
                                               _lib.ptr(d_feat.view), _lib.ptr(d_acc.view), _lib.ptr(d_view.view),
                                               _lib.ptr(g_basis.view), _lib.ptr(g_par.view), st), "shade_ref_bwd")
        torch.cuda.synchronize()
        for name, t in (("d_feat", d_feat), ("d_acc", d_acc), ("d_view", d_view), ("g_basis", g_basis), ("g_par", g_par)):
            assert t.intact(), (name, n)
            assert torch.isfinite(t.view).all(), (name, n)
        assert not (d_feat.view == SENT).any() and not (d_acc.view == SENT).any()
        assert float(g_par.view.abs().sum()) > 0
