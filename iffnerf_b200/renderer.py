"""`OctreeRender_trilinear_fast` with the reference's signature and return contract (renderer.py:12-25).

The reference slices the rays into `chunk`-sized pieces because its forward materialises
[chunk, S, 27] temporaries (459 MB at 4096 x 1036).  The fused kernels keep per-ray state in
registers, so the eval path launches over as many rays as the caller hands over (capped by
`tensorf.max_launch_rays`); CPU-resident rays are pipelined in slices (H2D copy stream + two alternating
compute streams, `tensorf.host_ray_slices`, 0 = auto).
`chunk` is honoured only where results could depend on it (the autograd / is_train path).
"""
from __future__ import annotations

import torch


def OctreeRender_trilinear_fast(rays, tensorf, chunk=4096, N_samples=-1, ndc_ray=False, bg_color=None, white_bg=None,
                                is_train=False, device="cuda", out_host=None):
    """Returns (rgb [N,3], None, depth [N], None, None) on `device`, like the reference.

    out_host (extension, eval path only): `(rgb_host [N,3], depth_host [N])` pinned CPU tensors that also receive the
    results — the download of each ray slice runs on its own stream behind the kernels of the next slice instead of
    after the whole call (what `evaluation()`'s `.cpu()` does in the reference, renderer.py:77-80).  The copies are
    ordered before anything the caller enqueues on the current stream afterwards."""
    if ndc_ray:
        raise NotImplementedError("ndc_ray sampling is outside the B200 render path")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("iffnerf_b200 renders on CUDA devices only (no CPU fallback); got device=%r" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if rays.is_cuda and rays.device != dev:
        rays = rays.to(dev)                      # outputs live on `device`, like the reference's per-chunk .to(device)
    n = rays.shape[0]
    grad_path = is_train or (torch.is_grad_enabled() and rays.requires_grad)
    if out_host is not None and (grad_path or len(out_host) != 2 or any(
            t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() for t in out_host)
            or out_host[0].shape != (n, 3) or out_host[1].shape != (n,)):
        raise ValueError("out_host must be (rgb [N,3], depth [N]) contiguous fp32 CPU tensors on the no-grad path")
    if grad_path:
        rgbs, depths = [], []
        for a in range(0, n, chunk):
            rgb, depth, _, _, _, _ = tensorf(rays[a:a + chunk].to(dev), is_train=is_train, bg_color=bg_color,
                                             white_bg=white_bg, ndc_ray=ndc_ray, N_samples=N_samples)
            rgbs.append(rgb)
            depths.append(depth)
        return torch.cat(rgbs), None, torch.cat(depths), None, None

    step = int(tensorf.max_launch_rays)
    rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((n,), dtype=torch.float32, device=dev)
    on_host = not rays.is_cuda
    if on_host and n > 0:
        # host-resident rays are cut into slices: the H2D copy of slice k+1 runs on a copy stream behind the kernels
        # of slice k, and consecutive slices launch on two alternating compute streams so the straggler CTAs at the
        # end of one slice's march overlap the start of the next instead of leaving SMs idle
        n_slices = int(getattr(tensorf, "host_ray_slices", 0)) or (2 if n >= (1 << 18) else 1)     # measured: 2 is best at 800x800
        step = min(step, max(1 << 14, -(-n // n_slices)))
    # slice boundaries; the FIRST slice is short (its upload is the only one nothing can hide behind), the rest of the
    # rays is cut into n_slices - 1 equal parts (each within max_launch_rays)
    bounds = list(range(0, n, step)) + [n]
    first_frac = float(getattr(tensorf, "host_first_slice_frac", 0.5))
    if on_host and len(bounds) > 2 and 0.0 < first_frac < 1.0:
        first = min(n, max(1 << 14, int(step * first_frac) // 4096 * 4096))
        rest_step = min(int(tensorf.max_launch_rays), max(1 << 14, -(-(n - first) // max(len(bounds) - 2, 1))))
        bounds = [0] + list(range(first, n, rest_step)) + [n]
    main = torch.cuda.current_stream(dev)
    if not on_host or n <= step:
        for a in range(0, n, step):
            cur = rays[a:a + step]
            if on_host:
                cur = cur.to(dev, non_blocking=True)
            tensorf.render_eval(cur, N_samples=N_samples, white_bg=bool(white_bg), bg_color=bg_color,
                                out_rgb=rgb[a:a + step], out_depth=depth[a:a + step])
        if out_host is not None:
            out_host[0].copy_(rgb, non_blocking=True)
            out_host[1].copy_(depth, non_blocking=True)
        return rgb, None, depth, None, None

    copy_stream, compute, down_stream = _side_streams(dev)
    tensorf.field_desc()                       # (re)pack parameter shadows on the caller's stream before forking
    tensorf._bg(bg_color, bool(white_bg), dev)  # ... and create the cached background constant there too
    if not tensorf.native_shade and tensorf.ref_kernel:
        tensorf.packed_ref_head()               # ... and the `Ref` head's packed parameters
    fork = torch.cuda.Event()
    fork.record(main)
    copy_stream.wait_event(fork)
    done = []
    for k, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
        with torch.cuda.stream(copy_stream):
            cur = rays[a:b].to(dev, non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(copy_stream)
        cs = compute[k % len(compute)]
        if k < len(compute):
            cs.wait_event(fork)
        cs.wait_event(arrived)
        with torch.cuda.stream(cs):
            cur.record_stream(cs)
            tensorf.render_eval(cur, N_samples=N_samples, white_bg=bool(white_bg), bg_color=bg_color,
                                out_rgb=rgb[a:b], out_depth=depth[a:b])
            ev = torch.cuda.Event()
            ev.record(cs)
        done.append(ev)
        if out_host is not None:            # this slice's download, behind the next slice's kernels
            down_stream.wait_event(ev)
            with torch.cuda.stream(down_stream):
                out_host[0][a:b].copy_(rgb[a:b], non_blocking=True)
                out_host[1][a:b].copy_(depth[a:b], non_blocking=True)
    for ev in done[-len(compute):]:
        main.wait_event(ev)
    if out_host is not None:
        rgb.record_stream(down_stream)
        depth.record_stream(down_stream)
        landed = torch.cuda.Event()
        landed.record(down_stream)
        main.wait_event(landed)
    return rgb, None, depth, None, None


_streams = {}


def _side_streams(dev):
    """(upload stream, [two compute streams], download stream) of a device, created once."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(device=dev), [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)],
                         torch.cuda.Stream(device=dev))
    return _streams[key]
