"""Autograd glue: the whole differentiable render as ONE torch.autograd.Function over the C ABI.

forward  = tvm_render_fwd (march kernel + fp32 shade kernel) -> rgb_map, depth_map, acc_map, alpha, z_vals, dists
backward = tvm_shade_bwd (d rgb_map -> d ray_feat, d acc, d viewdirs, d basis_mat, d MLP)
           -> tvm_march_bwd (re-march; vector-reduction scatter into the packed factor-gradient buffer, d rays)
           -> [optional data-parallel all-reduce of the packed buffer] -> tvm_unpack_factor_grads / tvm_unpack_mlp_grads

No torch op touches the render path in either direction; torch only owns the tensors and the autograd graph edge.
(`render_with_grad_torch_tail` keeps the earlier variant whose per-ray shading tail is torch autograd; it is used by
tests as an independent cross-check of tvm_shade_bwd.)
"""
from __future__ import annotations

import ctypes as C
import functools

import torch
import torch.nn.functional as F
from torch.autograd.function import once_differentiable

from . import _lib


def _guard_fwd(fn):
    """Runs a Function.forward(ctx, model, rays, ...) with the rays' GPU as the current device: the C-ABI launches go to
    the calling thread's current device, while streams and tensors belong to the rays' device."""
    @functools.wraps(fn)
    def wrapped(ctx, model, rays, *rest):
        if not rays.is_cuda:
            return fn(ctx, model, rays, *rest)           # raises the no-CPU-path error
        with torch.cuda.device(rays.device):
            return fn(ctx, model, rays, *rest)
    return wrapped


def _guard_bwd(fn):
    """Same for backward; also marks it once-differentiable (the kernels have no double backward: a
    loss.backward(create_graph=True) as in the reference's adahessian option must raise, not return zeros)."""
    @functools.wraps(fn)
    def wrapped(ctx, *grads):
        with torch.cuda.device(ctx.rays_c.device):
            return fn(ctx, *grads)
    return once_differentiable(wrapped)


def _param_state(params):
    return tuple((p.data_ptr(), p._version) for p in params)


def _check_unmodified(ctx):
    """The backward re-marches with the CURRENT parameters against the forward's workspace: refuse if a parameter
    changed in between (in-place optimiser step, .copy_, ...).  Edits through `.data` do not bump `_version`; call
    model.invalidate_packed() after those."""
    if _param_state(ctx.params) != ctx.param_state:
        raise _lib.TvmError("parameters were modified between forward and backward")


def _c(t):
    return None if t is None else t.detach().float().contiguous()


def _mlp_params(model):
    mods = [model.renderModule.mlp[i] for i in (0, 2, 4)]
    return [mods[0].weight, mods[0].bias, mods[1].weight, mods[1].bias, mods[2].weight, mods[2].bias]


def _split_flags(model, want_samples):
    """Differentiable forward without per-sample outputs (the pose loop): the two-kernel march of the eval path.  The
    backward kernels read every ray's ray_feat row, so rays without appearance samples get their zero rows."""
    if model.split_app and not want_samples:
        return _lib.F_SPLIT_APP | _lib.F_ZERO_UNLIT
    return 0


class _Render(torch.autograd.Function):
    """inputs: model, rays, S, jitter, bg, flags, 12 factors, basis, w1, b1, w2, b2, w3, b3.
    flags & F_EARLY_TERM: the caller does not want the per-sample outputs (alpha / z_vals / dists come back as None),
    so forward and backward both stop a ray at T < early_term_eps."""

    @staticmethod
    @_guard_fwd
    def forward(ctx, model, rays, S, jitter, bg, flags, *params):
        from .tensorf import _stream
        rays_c = model._prep_rays(rays)
        dev = rays_c.device
        n = rays_c.shape[0]
        d, keep = model.field_desc()
        lib = _lib.load()
        want_samples = not (flags & _lib.F_EARLY_TERM)
        fwd_flags = flags | _split_flags(model, want_samples)
        need = C.c_size_t(0)
        _lib.check(lib.tvm_workspace_bytes(C.byref(d), n, fwd_flags & _lib.F_SPLIT_APP, C.byref(need)),
                   "tvm_workspace_bytes")
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        rgb = torch.empty((n, 3), device=dev)
        depth = torch.empty((n,), device=dev)
        acc = torch.empty((n,), device=dev)
        alpha, z, dists = (torch.empty((n, S), device=dev) for _ in range(3)) if want_samples else (None, None, None)
        jit = None if jitter is None else jitter.detach().to(dev).float().reshape(-1).contiguous()
        bg_c = _c(bg)
        _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], S, _lib.ptr(jit),
                                      _lib.ptr(bg_c), fwd_flags, _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(acc),
                                      _lib.ptr(alpha), _lib.ptr(z), _lib.ptr(dists), None, None, None,
                                      _lib.ptr(ws), ws.numel(), _stream(dev)), "tvm_render_fwd")
        ctx.model, ctx.S, ctx.jit, ctx.rays_c, ctx.ws, ctx.bg = model, S, jit, rays_c, ws, bg_c
        ctx.ray_cols = rays.shape[1]
        ctx.flags = flags & ~(_lib.F_MLP_TC3 | _lib.F_MLP_BF16)      # the backward kernels take the sampler bits only
        ctx.params, ctx.param_state = params, _param_state(params)
        if want_samples:
            ctx.mark_non_differentiable(depth, z, dists)
        else:
            ctx.mark_non_differentiable(depth)
        return rgb, depth, acc, alpha, z, dists

    @staticmethod
    @_guard_bwd
    def backward(ctx, g_rgb, g_depth, g_acc, g_alpha, g_z, g_dists):
        from .tensorf import _stream
        model, rays_c = ctx.model, ctx.rays_c
        dev = rays_c.device
        n = rays_c.shape[0]
        need = ctx.needs_input_grad
        want_rays = need[1]
        want_factors = any(need[6:18])
        want_basis = need[18]
        want_mlp = any(need[19:25])
        _check_unmodified(ctx)
        d, keep = model.field_desc()
        lib = _lib.load()
        st = _stream(dev)
        ta = sum(model.app_n_comp)
        g_rgb = _c(g_rgb) if g_rgb is not None else torch.zeros((n, 3), device=dev)
        g_acc, g_alpha = _c(g_acc), _c(g_alpha)
        d_feat = torch.empty((n, ta), device=dev)
        d_acc = torch.empty((n,), device=dev)
        d_view = torch.empty((n, 3), device=dev) if want_rays else None
        # parameter gradients accumulate in the model's persistent, pre-zeroed workspace [factors | MLP | basis]
        want_params = want_factors or want_basis or want_mlp
        ws_all = ws_f = ws_m = ws_b = None
        if want_params:
            ws_all, ws_f, ws_m, ws_b = model.grad_workspace(dev)
        g_basis_buf = ws_b if want_basis else None
        g_mlp = ws_m if want_mlp else None
        _lib.check(lib.tvm_shade_bwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], _lib.ptr(ctx.bg),
                                     _lib.ptr(g_rgb), _lib.ptr(g_acc), _lib.ptr(d_feat), _lib.ptr(d_acc),
                                     _lib.ptr(g_basis_buf), _lib.ptr(g_mlp), _lib.ptr(d_view), _lib.ptr(ctx.ws),
                                     ctx.ws.numel(), st), "tvm_shade_bwd")
        g_packed = ws_f if want_factors else None
        g_rays6 = torch.zeros((n, 6), device=dev) if want_rays else None
        _lib.check(lib.tvm_march_bwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], ctx.S, _lib.ptr(ctx.jit),
                                     ctx.flags, _lib.ptr(d_feat), _lib.ptr(d_acc), _lib.ptr(g_alpha), _lib.ptr(g_packed),
                                     _lib.ptr(g_rays6), _lib.ptr(ctx.ws), ctx.ws.numel(), st), "tvm_march_bwd")
        scale = 1.0
        sync = getattr(model, "grad_sync", None)
        if sync is not None and want_params:
            # data parallel: ONE all-reduce of the whole workspace (factors + MLP + basis), on this stream, before
            # unpacking; the 1/world of the mean rides on the unpack
            scale = sync.reduce_sum(ws_all, covers_small_params=True)
        factor_grads = [None] * 12
        if want_factors:
            planes, lines = model._factor_params()
            gp = [torch.empty_like(p) for p in planes]
            gl = [torch.empty_like(p) for p in lines]
            _lib.check(lib.tvm_unpack_factor_grads_scaled(C.byref(d), _lib.ptr(g_packed), _lib.ptr_array(gp),
                                                          _lib.ptr_array(gl), 0, scale, 1, st), "tvm_unpack_factor_grads")
            factor_grads = gp + gl
        mlp_grads = [None] * 6
        g_basis = None
        if want_mlp or want_basis:
            small = ws_all[ws_f.numel():]
            if scale != 1.0:
                small.mul_(scale)
            if want_mlp:
                mlp_grads = [torch.empty_like(p) for p in _mlp_params(model)]
                _lib.check(lib.tvm_unpack_mlp_grads(C.byref(d), _lib.ptr(g_mlp), *[_lib.ptr(t) for t in mlp_grads], 0, st),
                           "tvm_unpack_mlp_grads")
            if want_basis:
                g_basis = ws_b.clone()
            small.zero_()
        d_rays = None
        if want_rays:
            d_rays = torch.zeros((n, ctx.ray_cols), device=dev)
            d_rays[:, :6] = g_rays6
            d_rays[:, 3:6] += d_view
        return (None, d_rays, None, None, None, None, *factor_grads, g_basis, *mlp_grads)


def render_with_grad(model, rays_chunk, white_bg, bg_color, N_samples, jitter, point_samples=False, want_samples=True):
    """Differentiable TensorBase.forward (models/tensorBase.py:775-917): 6-tuple, `depth_map` without grad.
    want_samples=False (eval only): alpha / z_vals / dists are None and rays terminate early in both directions."""
    S = N_samples if N_samples > 0 else model.nSamples
    planes, lines = model._factor_params()
    rays = rays_chunk if rays_chunk.dtype == torch.float32 else rays_chunk.float()
    bg = model._bg(bg_color, white_bg, rays.device)
    flags = _lib.F_POINT_SAMPLES if point_samples else 0
    if not want_samples and model.early_term_eps > 0:
        flags |= _lib.F_EARLY_TERM
    if getattr(model, "bwd_runs", False):
        flags |= _lib.F_BWD_RUNS                # opt-in: run-aggregated gradient scatter (see csrc/march_bwd.cu)
    if (model.grad_forward_tc3 and model.native_shade and model._shade_mode() == "tc3"
            and not any(p.requires_grad for p in _mlp_params(model))):
        # frozen head (pose refinement): the forward shades on the tensor cores (bf16x3 split, 5e-7 from the FFMA
        # kernel); tvm_shade_bwd re-derives the activations in fp32 either way
        flags |= _lib.F_MLP_TC3
    return _Render.apply(model, rays, S, jitter, bg, flags, *planes, *lines, model.basis_mat.weight,
                         *_mlp_params(model))


# ----------------------------------------------------------------------------------------------------------------
# march stage through the C ABI, shading tail as torch autograd: the path of the `Ref` head (no fused kernel yet) and
# an independent cross-check of tvm_shade_bwd for MLP_Fea
# ----------------------------------------------------------------------------------------------------------------
class _March(torch.autograd.Function):
    @staticmethod
    @_guard_fwd
    def forward(ctx, model, rays, S, jitter, flags, *factors):
        from .tensorf import _stream
        rays_c = model._prep_rays(rays)
        dev = rays_c.device
        n = rays_c.shape[0]
        d, keep = model.field_desc()
        lib = _lib.load()
        want_samples = not (flags & _lib.F_EARLY_TERM)     # without per-sample outputs both directions terminate early
        fwd_flags = flags | _split_flags(model, want_samples)
        need = C.c_size_t(0)
        _lib.check(lib.tvm_workspace_bytes(C.byref(d), n, fwd_flags & _lib.F_SPLIT_APP, C.byref(need)),
                   "tvm_workspace_bytes")
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        alpha, z, dists = (torch.empty((n, S), device=dev) for _ in range(3)) if want_samples else (None, None, None)
        jit = None if jitter is None else jitter.detach().to(dev).float().reshape(-1).contiguous()
        bg = model._bg(None, False, dev)
        _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], S, _lib.ptr(jit),
                                      _lib.ptr(bg), _lib.F_NO_SHADE | fwd_flags, None, None, None, _lib.ptr(alpha), _lib.ptr(z),
                                      _lib.ptr(dists), None, None, None, _lib.ptr(ws), ws.numel(), _stream(dev)),
                   "tvm_render_fwd")
        v = model.workspace_views(d, ws, n)
        ctx.model, ctx.S, ctx.jit, ctx.rays_c, ctx.ws, ctx.flags = model, S, jit, rays_c, ws, flags
        ctx.ray_cols = rays.shape[1]
        ctx.params, ctx.param_state = factors, _param_state(factors)
        if want_samples:
            ctx.mark_non_differentiable(v["depth"], z, dists, v["app_count"])
        else:
            ctx.mark_non_differentiable(v["depth"], v["app_count"])
        return v["ray_feat"], v["acc"], v["depth"], alpha, z, dists, v["app_count"]

    @staticmethod
    @_guard_bwd
    def backward(ctx, g_feat, g_acc, g_depth, g_alpha, g_z, g_dists, g_cnt):
        from .tensorf import _stream
        model, rays_c = ctx.model, ctx.rays_c
        dev = rays_c.device
        n = rays_c.shape[0]
        want_rays = ctx.needs_input_grad[1]
        want_factors = any(ctx.needs_input_grad[5:])
        _check_unmodified(ctx)
        d, keep = model.field_desc()
        lib = _lib.load()
        g_packed = torch.zeros(int(d.n_factor_floats), device=dev) if want_factors else None
        g_rays = torch.zeros((n, 6), device=dev) if want_rays else None
        _lib.check(lib.tvm_march_bwd(C.byref(d), _lib.ptr(rays_c), n, rays_c.shape[1], ctx.S, _lib.ptr(ctx.jit),
                                     ctx.flags, _lib.ptr(_c(g_feat)), _lib.ptr(_c(g_acc)), _lib.ptr(_c(g_alpha)),
                                     _lib.ptr(g_packed), _lib.ptr(g_rays), _lib.ptr(ctx.ws), ctx.ws.numel(),
                                     _stream(dev)), "tvm_march_bwd")
        grads = [None] * 12
        if want_factors:
            sync = getattr(model, "grad_sync", None)
            if sync is not None:
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
        return (None, d_rays, None, None, None, *grads)


class _RefTail(torch.autograd.Function):
    """Fused `Ref` tail under autograd: forward = tvm_shade_ref_fwd on the march outputs, backward = tvm_shade_ref_bwd.
    inputs: model, rays, bg, ray_feat, acc, depth_partial, app_count, basis, then the 12 head tensors
    (normal / tint / rough / diffuse / bottleneck / specular: weight, bias each)."""

    NAMES = ("normal", "tint", "rough", "diffuse", "bott", "spec")

    @staticmethod
    def _lins(model):
        rm = model.renderModule
        return [rm.normal_mlp[0], rm.tint_color_mlp[0], rm.roughness_mlp[0], rm.diffuse_color_mlp[0], rm.bottleneck_mlp,
                rm.specular_mlp[0]]

    @staticmethod
    @_guard_fwd
    def forward(ctx, model, rays, bg, ray_feat, acc, depth_p, app_count, basis, *head):
        from .tensorf import _stream
        dev = ray_feat.device
        rays_c = model._prep_rays(rays)
        n = rays_c.shape[0]
        d, keep = model.field_desc()
        h, buf = model.packed_ref_head()
        lib = _lib.load()
        # the march outputs in the workspace layout tvm_shade_ref_fwd reads
        need = C.c_size_t(0)
        _lib.check(lib.tvm_workspace_bytes(C.byref(d), n, 0, C.byref(need)), "tvm_workspace_bytes")
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        v = model.workspace_views(d, ws, n)
        v["ray_feat"].copy_(ray_feat); v["acc"].copy_(acc); v["depth"].copy_(depth_p); v["app_count"].copy_(app_count)
        rgb = torch.empty((n, 3), device=dev)
        depth = torch.empty((n,), device=dev)
        bg_c = _c(bg)
        _lib.check(lib.tvm_shade_ref_fwd(C.byref(d), C.byref(h), _lib.ptr(rays_c), n, rays_c.shape[1], _lib.ptr(bg_c),
                                         _lib.ptr(rgb), _lib.ptr(depth), None, _lib.ptr(ws), ws.numel(), _stream(dev)),
                   "tvm_shade_ref_fwd")
        ctx.model, ctx.rays_c, ctx.bg, ctx.ray_cols = model, rays_c, bg_c, rays.shape[1]
        ctx.save_for_backward(ray_feat, acc, app_count)
        ctx.params = (basis,) + tuple(head)
        ctx.param_state = _param_state(ctx.params)
        ctx.mark_non_differentiable(depth)
        return rgb, depth

    @staticmethod
    @_guard_bwd
    def backward(ctx, g_rgb, g_depth):
        from .tensorf import _stream
        model, rays_c = ctx.model, ctx.rays_c
        ray_feat, acc, app_count = ctx.saved_tensors
        dev = ray_feat.device
        n, ta = ray_feat.shape
        need = ctx.needs_input_grad
        _check_unmodified(ctx)
        d, keep = model.field_desc()
        h, buf = model.packed_ref_head()
        lib = _lib.load()
        want_rays, want_basis, want_head = need[1], need[7], any(need[8:])
        d_feat = torch.empty((n, ta), device=dev)
        d_acc = torch.empty((n,), device=dev)
        d_view = torch.empty((n, 3), device=dev) if want_rays else None
        g_basis = torch.zeros_like(model.basis_mat.weight) if want_basis else None
        g_par = torch.zeros_like(buf) if want_head else None
        g_rgb_c = _c(g_rgb) if g_rgb is not None else torch.zeros((n, 3), device=dev)
        _lib.check(lib.tvm_shade_ref_bwd(C.byref(d), C.byref(h), _lib.ptr(rays_c), n, rays_c.shape[1], _lib.ptr(ctx.bg),
                                         _lib.ptr(_c(ray_feat)), _lib.ptr(_c(acc)), _lib.ptr(app_count.contiguous()),
                                         _lib.ptr(g_rgb_c), None, _lib.ptr(d_feat), _lib.ptr(d_acc), _lib.ptr(d_view),
                                         _lib.ptr(g_basis), _lib.ptr(g_par), _stream(dev)), "tvm_shade_ref_bwd")
        head_grads = [None] * 12
        if want_head:
            offs = (C.c_int32 * 8)()
            _lib.check(lib.tvm_ref_head_layout(C.byref(h), offs), "tvm_ref_head_layout")
            small_w, bott_w, small_b, bott_b, spec_w, spec_b, ide, in4 = list(offs)
            in_c, fc = h.in_c, h.feature_c
            sw = g_par[small_w:small_w + 10 * in4].view(10, in4)[:, :in_c]
            sb = g_par[small_b:small_b + 10]
            rows = {"normal": (0, 3), "tint": (3, 6), "rough": (6, 7), "diffuse": (7, 10)}
            lins = _RefTail._lins(model)
            out = []
            for name, lin in zip(_RefTail.NAMES, lins):
                if name in rows:
                    a_, b_ = rows[name]
                    out += [sw[a_:b_].contiguous(), sb[a_:b_].contiguous()]
                elif name == "bott":
                    out += [g_par[bott_w:bott_w + fc * in4].view(fc, in4)[:, :in_c].contiguous(), g_par[bott_b:bott_b + fc].contiguous()]
                else:
                    nin = lin.in_features
                    out += [g_par[spec_w:spec_w + 3 * nin].view(3, nin).contiguous(), g_par[spec_b:spec_b + 3].contiguous()]
            head_grads = out
        d_rays = None
        if want_rays:
            d_rays = torch.zeros((n, ctx.ray_cols), device=dev)
            d_rays[:, 3:6] = d_view
        return (None, d_rays, None, d_feat, d_acc, None, None, g_basis, *head_grads)


def _ref_tail_params(model):
    ps = []
    for lin in _RefTail._lins(model):
        ps += [lin.weight, lin.bias]
    return ps


def render_with_grad_torch_tail(model, rays_chunk, white_bg, bg_color, N_samples, jitter, point_samples=False,
                                want_samples=True):
    S = N_samples if N_samples > 0 else model.nSamples
    planes, lines = model._factor_params()
    rays = rays_chunk if rays_chunk.dtype == torch.float32 else rays_chunk.float()
    flags = _lib.F_POINT_SAMPLES if point_samples else 0
    if not want_samples and model.early_term_eps > 0:
        flags |= _lib.F_EARLY_TERM
    ray_feat, acc, depth_p, alpha, z, dists, app_count = _March.apply(model, rays, S, jitter, flags, *planes, *lines)
    if bg_color is None:
        bg_color = model._bg(None, white_bg, rays.device)
    if getattr(model, "ref_kernel_train", False) and not model.native_shade and model.packed_ref_head() is None:
        raise _lib.TvmError("this `Ref` head configuration has no fused tail kernel and the render path has no eager "
                            "fallback; ref_kernel_train = False runs the torch-op tail explicitly (cross-checks only)")
    if getattr(model, "ref_kernel_train", False) and not model.native_shade:
        # fused Ref tail in both directions (csrc/shade_ref.cu); d(rgb_map)/d(acc) reaches the march backward through acc
        rgb_map, depth_map = _RefTail.apply(model, rays, bg_color, ray_feat, acc, depth_p, app_count,
                                            model.basis_mat.weight, *_ref_tail_params(model))
        return rgb_map, depth_map, acc, alpha, z, dists
    view = rays[:, 3:6]
    feat = F.linear(ray_feat, model.basis_mat.weight)
    rgb, _ = model.renderModule(None, view, feat, None)
    rgb = rgb * (app_count > 0).to(rgb.dtype)[:, None]
    if bg_color is None:
        bg_color = model._bg(None, white_bg, rays.device)
    rgb_map = (rgb * acc[..., None] + bg_color * (1.0 - acc[..., None])).clamp(0, 1)
    with torch.no_grad():
        depth_map = depth_p + (1.0 - acc) * rays[..., -1]
    return rgb_map, depth_map, acc, alpha, z, dists
