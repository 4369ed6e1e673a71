"""Autograd glue: the march stage as a torch.autograd.Function over the C ABI.

forward  = tvm_render_fwd(TVM_F_NO_SHADE) -> ray_feat [N, sum(n_app)], acc [N], depth partial [N],
           alpha / z_vals / dists [N,S]
backward = tvm_march_bwd (re-march + vector-reduction scatter) -> gradients of the 12 factor tensors in the
           reference [1,C,H,W] layout and, if the rays require grad (pose refinement), d(rays).

The per-ray shading tail (basis_mat, MLPRender_Fea, background blend; models/tensorBase.py:886-904) runs as
torch ops ON THE GPU in this differentiable path so that autograd provides d(basis_mat), d(MLP) and
d(viewdirs); the eval path uses the fused shade kernel instead.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib


class _March(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, rays, S, jitter, *factors):
        from .tensorf import _stream
        rays_c = model._prep_rays(rays)
        dev = rays_c.device
        n = rays_c.shape[0]
        d, keep = model.field_desc()
        lib = _lib.load()
        need = C.c_size_t(0)
        _lib.check(lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need)), "tvm_workspace_bytes")
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        alpha, z, dists = (torch.empty((n, S), device=dev) for _ in range(3))
        jit = None if jitter is None else jitter.detach().to(dev).float().reshape(-1).contiguous()
        bg = model._bg(None, False, dev)
        _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], S, _lib.ptr(jit),
                                      _lib.ptr(bg), _lib.F_NO_SHADE, None, None, None, _lib.ptr(alpha), _lib.ptr(z),
                                      _lib.ptr(dists), None, None, None, _lib.ptr(ws), ws.numel(), _stream(dev)),
                   "tvm_render_fwd")
        v = model.workspace_views(d, ws, n)
        ctx.model, ctx.S, ctx.jit, ctx.rays_c, ctx.ws = model, S, jit, rays_c, ws
        ctx.ray_cols = rays.shape[1]
        ctx.packed_key = model._packed_key
        ctx.mark_non_differentiable(v["depth"], z, dists, v["app_count"])
        return v["ray_feat"], v["acc"], v["depth"], alpha, z, dists, v["app_count"]

    @staticmethod
    def backward(ctx, g_feat, g_acc, g_depth, g_alpha, g_z, g_dists, g_cnt):
        from .tensorf import _stream
        model, rays_c = ctx.model, ctx.rays_c
        dev = rays_c.device
        n = rays_c.shape[0]
        want_rays = ctx.needs_input_grad[1]
        want_factors = any(ctx.needs_input_grad[4:])
        if model._packed_key != ctx.packed_key:
            raise _lib.TvmError("factor parameters were modified between forward and backward")
        d, keep = model.field_desc()
        lib = _lib.load()
        g_packed = torch.zeros(int(d.n_factor_floats), device=dev) if want_factors else None
        g_rays = torch.zeros((n, 6), device=dev) if want_rays else None

        def c(t):
            return None if t is None else t.detach().float().contiguous()
        g_feat, g_acc, g_alpha = c(g_feat), c(g_acc), c(g_alpha)
        _lib.check(lib.tvm_march_bwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], ctx.S, _lib.ptr(ctx.jit),
                                     _lib.ptr(g_feat), _lib.ptr(g_acc), _lib.ptr(g_alpha), _lib.ptr(g_packed),
                                     _lib.ptr(g_rays), _lib.ptr(ctx.ws), ctx.ws.numel(), _stream(dev)),
                   "tvm_march_bwd")
        grads = [None] * 12
        if want_factors:
            sync = getattr(model, "grad_sync", None)
            if sync is not None:
                # data parallel: ONE all-reduce of the flat packed buffer, on this stream, before unpacking
                sync.reduce_packed_factor_grads(g_packed)
            planes, lines = model._factor_params()
            gp = [torch.empty_like(p) for p in planes]
            gl = [torch.empty_like(p) for p in lines]
            _lib.check(lib.tvm_unpack_factor_grads(C.byref(d), _lib.ptr(g_packed), _lib.ptr_array(gp),
                                                   _lib.ptr_array(gl), 0, _stream(dev)), "tvm_unpack_factor_grads")
            grads = gp + gl
        d_rays = None
        if want_rays:
            d_rays = torch.zeros((n, ctx.ray_cols), device=dev)
            d_rays[:, :6] = g_rays
        return (None, d_rays, None, None, *grads)


def render_with_grad(model, rays_chunk, white_bg, bg_color, N_samples, jitter):
    """Differentiable TensorBase.forward (models/tensorBase.py:775-917): 6-tuple, `depth_map` without grad."""
    S = N_samples if N_samples > 0 else model.nSamples
    planes, lines = model._factor_params()
    rays = rays_chunk if rays_chunk.dtype == torch.float32 else rays_chunk.float()
    ray_feat, acc, depth_p, alpha, z, dists, app_count = _March.apply(model, rays, S, jitter, *planes, *lines)
    view = rays[:, 3:6]
    feat = F.linear(ray_feat, model.basis_mat.weight)
    rgb, _ = model.renderModule(None, view, feat, None)
    rgb = rgb * (app_count > 0).to(rgb.dtype)[:, None]            # rays_to_consider (:886-896), no host sync
    if bg_color is None:
        bg_color = model._bg(None, white_bg, rays.device)
    rgb_map = (rgb * acc[..., None] + bg_color * (1.0 - acc[..., None])).clamp(0, 1)
    with torch.no_grad():
        depth_map = depth_p + (1.0 - acc) * rays[..., -1]
    return rgb_map, depth_map, acc, alpha, z, dists
