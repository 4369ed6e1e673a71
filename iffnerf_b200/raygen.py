"""Fused pixel -> ray generation with the pose gradient.

The reference's iNeRF step rebuilds FIVE full-image [H,W,3] grids from the current pose and then indexes the
1024 pixels it renders (inerf/estimate_pose_inerf.py:149-164 over ray_utils.py:28-100).  `pixel_rays` builds
exactly the rays that are rendered — `[N,7] = (origin, direction, radii)` — in one launch, and its backward
reduces d(rays) straight into the 3x4 pose gradient, so the autograd chain
`CameraTransfer -> c2w -> rays -> TensorVMSplit.forward` keeps working with two tiny kernels in place of the grids.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _lib

NORMALIZE_VIEWDIRS = 1     # directions = ori / |ori| before the rotation (blender.py:70-72, estimate_pose_inerf.py:97)
RENORMALIZE = 2            # F.normalize(rays_d) after it (estimate_pose_inerf.py:159)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


_kinv_cache = {}


def _kinv(K):
    """Host float[9] of torch.inverse(K) (ray_utils.py:50); K is [3,3] or [1,3,3].  Cached (CPU intrinsics by value,
    device intrinsics by tensor identity + version): the intrinsics are constants of a run, and a cached value keeps
    `pixel_rays` free of host<->device traffic (CUDA-graph capturable) even when K lives on the GPU."""
    if K.device.type == "cpu":
        key = ("cpu",) + tuple(K.detach().reshape(-1)[:9].tolist())     # by value: no aliasing of recycled addresses
        keep = None
    else:
        key = ("dev", K.data_ptr(), K._version)                          # by identity; the entry keeps K alive, so the
        keep = K                                                         # address cannot be recycled while it is cached
    hit = _kinv_cache.get(key)
    if hit is None:
        k = torch.inverse(K.detach().reshape(-1, 3, 3)[0].float().cpu()).contiguous()
        hit = ((C.c_float * 9)(*k.reshape(-1).tolist()), keep)
        if len(_kinv_cache) > 64:
            _kinv_cache.clear()
        _kinv_cache[key] = hit
    return hit[0]


class _PixelRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c2w, kinv, pixels, pose_index, width, n, flags):
        if not c2w.is_cuda:
            raise _lib.TvmError("pixel_rays needs the pose on a CUDA device (there is no CPU path)")
        dev = c2w.device
        m = c2w.detach().float().reshape(-1, c2w.shape[-2], 4)
        if m.shape[1] not in (3, 4):
            raise ValueError(f"c2w must be [..,3|4,4], got {tuple(c2w.shape)}")
        m = m.contiguous()
        rays = torch.empty((n, 7), dtype=torch.float32, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            _lib.check(lib.tvm_pixel_rays_fwd(_lib.ptr(m), m.shape[1] * 4, kinv, _lib.ptr(pixels), _lib.ptr(pose_index),
                                              width, n, flags, _lib.ptr(rays), _stream(dev)), "tvm_pixel_rays_fwd")
        ctx.save_for_backward(m, pixels, pose_index)
        ctx.meta = (kinv, width, n, flags, tuple(c2w.shape))
        return rays

    @staticmethod
    @once_differentiable
    def backward(ctx, g_rays):
        m, pixels, pose_index = ctx.saved_tensors
        kinv, width, n, flags, shape = ctx.meta
        g = g_rays.detach().float().contiguous()
        P = m.shape[0]
        g_c2w = torch.zeros((P, 3, 4), dtype=torch.float32, device=m.device)
        lib = _lib.load()
        with torch.cuda.device(m.device):
            _lib.check(lib.tvm_pixel_rays_bwd(_lib.ptr(m), m.shape[1] * 4, kinv, _lib.ptr(pixels), _lib.ptr(pose_index),
                                              width, n, flags, _lib.ptr(g), g.shape[1], _lib.ptr(g_c2w),
                                              _stream(m.device)), "tvm_pixel_rays_bwd")
        if m.shape[1] == 4:
            g_c2w = torch.cat([g_c2w, torch.zeros((P, 1, 4), device=m.device)], dim=1)
        return g_c2w.reshape(shape), None, None, None, None, None, None


def pixel_rays(K, c2w, pixels=None, pose_index=None, image_wh=None, renormalize=True):
    """Rays `[N,7] = (o, d, radii)` of `pixels` [N,2] (x, y) seen through pose(s) `c2w` ([3|4,4] or [P,3|4,4]).

    `pose_index` [N] selects the pose of each ray when several candidate poses are batched (BASELINE config 5);
    `pixels=None` with `image_wh=(W, H)` generates the full image in row-major order (the loaders' layout,
    dataLoader/blender.py:105-114; pass `renormalize=False` for their exact semantics).  Differentiable w.r.t. `c2w`."""
    dev = c2w.device
    if pixels is not None:
        pixels = pixels.to(device=dev, dtype=torch.int32).contiguous()
        n, width = int(pixels.shape[0]), 0
    else:
        if image_wh is None:
            raise ValueError("give either pixels or image_wh")
        width = int(image_wh[0])
        n_pose = 1 if c2w.dim() == 2 else int(c2w.shape[0])
        if n_pose != 1 and pose_index is None:
            raise ValueError("full-image generation takes one pose (or an explicit pose_index)")
        n = width * int(image_wh[1])
    if pose_index is not None:
        if not pose_index.is_cuda and pose_index.numel() > 0:      # host-side indices are range-checked for free
            n_pose = 1 if c2w.dim() == 2 else int(c2w.shape[0])
            if int(pose_index.min()) < 0 or int(pose_index.max()) >= n_pose:
                raise ValueError(f"pose_index out of range for {n_pose} pose(s)")
        pose_index = pose_index.to(device=dev, dtype=torch.int32).contiguous()
    flags = NORMALIZE_VIEWDIRS | (RENORMALIZE if renormalize else 0)
    return _PixelRays.apply(c2w, _kinv(K), pixels, pose_index, width, n, flags)


def pixel_rays_lie(K, c2w_se3, pixels=None, pose_index=None, image_wh=None, renormalize=True):
    """`get_rays_lie` (ray_utils.py:103-140; no caller in the reference): the same rays from a pose given as a Lie-group
    element — any object with `.rotation.matrix()` ([...,3,3]) and `.t` ([...,3]) like kornia's `Se3`.  The 3x4 pose
    is assembled with torch ops, so gradients flow back into the group parameters through `pixel_rays`."""
    rot = c2w_se3.rotation.matrix()
    c2w = torch.cat([rot, c2w_se3.t.unsqueeze(-1)], dim=-1)
    return pixel_rays(K, c2w, pixels=pixels, pose_index=pose_index, image_wh=image_wh, renormalize=renormalize)
