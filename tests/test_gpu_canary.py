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
        self.buf = torch.full((n + 2 * PAD,), SENT, dtype=dtype, device=dev)
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
