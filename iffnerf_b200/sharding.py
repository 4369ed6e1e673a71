"""Ray-sharded data parallelism across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-GPU; this is the B200-native addition of SURVEY.md 8e:
  * factors + MLP are replicated (69.5 MB at 300^3), rays are the sharded unit;
  * forward/eval needs NO data-path collective — each rank renders its tiles and the results are
    written back at their original offsets (optionally all-gathered, 16 B/ray);
  * tiles are dealt round-robin (cyclic) because contiguous image rows are badly load-imbalanced
    (background rays leave the field at the occupancy test, object rays march hundreds of samples);
  * training all-reduces ONE flat fp32 bucket — the packed factor-gradient buffer the backward kernel
    scatters into — over NCCL/NVLink, on the stream the kernel ran on, before it is unpacked into the
    reference-layout .grad tensors; basis/MLP gradients (90 KB) go in a second small bucket.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

DEFAULT_TILE = 4096


def shard_tiles(n_rays: int, world: int, rank: int, tile: int = DEFAULT_TILE) -> List[Tuple[int, int]]:
    """[start, stop) ray ranges owned by `rank`: tile t belongs to rank t % world."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    n_tiles = (n_rays + tile - 1) // tile
    return [(t * tile, min((t + 1) * tile, n_rays)) for t in range(rank, n_tiles, world)]


_index_cache = {}
_peer_fallback_reason = None      # why "auto" placement fell back to the NCCL gather (None: it did not)


def shard_index(n_rays: int, world: int, rank: int, tile: int = DEFAULT_TILE, device=None) -> torch.Tensor:
    """Flat ray indices owned by `rank` (concatenation of its tiles, ascending).  Built with a handful of vectorised ops
    (local position j -> ((j // tile) * world + rank) * tile + j % tile) and cached per (n, world, rank, tile, device):
    a 1080p image has 507 tiles, and one `arange` launch per tile made the sharded render host-bound."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    key = (n_rays, world, rank, tile, str(device))
    hit = _index_cache.get(key)
    if hit is None:
        n_tiles = (n_rays + tile - 1) // tile
        mine = len(range(rank, n_tiles, world))
        j = torch.arange(mine * tile, device=device)
        g = (torch.div(j, tile, rounding_mode="floor") * world + rank) * tile + j % tile
        hit = g[g < n_rays]
        if len(_index_cache) > 64:
            _index_cache.clear()
        _index_cache[key] = hit
    return hit


def local_rays(rays: torch.Tensor, world: int, rank: int, tile: int = DEFAULT_TILE) -> Tuple[torch.Tensor, torch.Tensor]:
    idx = shard_index(rays.shape[0], world, rank, tile, device=rays.device)
    return rays.index_select(0, idx), idx


class PeerImages:
    """Image buffers (rgb [N,3] + depth [N]) of every rank of a group in NVLink peer memory (torch symmetric memory):
    `rgb_ptrs[k]` / `depth_ptrs[k]` are rank k's buffers as seen from THIS GPU, so a kernel here can store into all of
    them.  Double-buffered: a frame's result stays valid until the second-next frame of the same size starts."""

    def __init__(self, n, device, group=None):
        import torch.distributed._symmetric_memory as symm
        pg = group if group is not None else dist.group.WORLD
        self.n, self.frames, self.flip = n, [], 0
        for _ in range(2):
            buf = symm.empty((4 * n,), dtype=torch.float32, device=device)
            hdl = symm.rendezvous(buf, pg.group_name)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            self.frames.append({"buf": buf, "hdl": hdl, "rgb": buf[:3 * n].view(n, 3), "depth": buf[3 * n:],
                                "rgb_ptrs": ptrs, "depth_ptrs": [p + 12 * n for p in ptrs]})

    def next_frame(self):
        self.flip ^= 1
        return self.frames[self.flip]


_peer_cache = {}


def peer_images(n, device, group=None):
    key = (n, str(device), id(group))
    if key not in _peer_cache:
        if len(_peer_cache) > 8:
            _peer_cache.clear()
        _peer_cache[key] = PeerImages(n, device, group)
    return _peer_cache[key]


def render_sharded_peer(rays, tensorf, group=None, tile: int = DEFAULT_TILE, device=None, **render_kw):
    """Ray-sharded render whose shading epilogue IS the all-gather: every rank renders its cyclic tiles and its shading
    kernel stores each ray's rgb / depth straight into the image buffers of ALL ranks at the ray's original offset
    (NVLink peer stores through torch symmetric memory, tvm_scatter_out); one device-side barrier publishes the frame.
    No index_put / pad / all_gather / permute.  Returns (rgb [N,3], depth [N]) views of this rank's buffer; they stay
    valid until the second-next call with the same N (double buffering)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = rays.shape[0]
    dev = torch.device(device) if device is not None else rays.device
    mine, idx = local_rays(rays.to(dev) if rays.device != dev else rays, world, rank, tile)
    frame = peer_images(n, dev, group).next_frame()
    kw = {k: v for k, v in render_kw.items() if k in ("N_samples", "white_bg", "bg_color")}
    tensorf.render_eval(mine, scatter=(idx, frame["rgb_ptrs"], frame["depth_ptrs"]), **kw)
    frame["hdl"].barrier(channel=0)          # on the current stream: every rank's stores have landed when it completes
    return frame["rgb"], frame["depth"]


def render_sharded(rays, tensorf, renderer, group=None, tile: int = DEFAULT_TILE, gather: bool = True,
                   device=None, placement: str = "auto", **render_kw):
    """Renders this rank's cyclic tiles of `rays` with `renderer` (OctreeRender_trilinear_fast signature) and
    returns (rgb [N,3], depth [N]) for ALL rays when `gather` (every rank gets the full image), else the local
    slice plus its ray indices.  No collective touches the render itself.

    placement (with gather): "peer" = the shading kernel writes into every rank's image over NVLink peer memory
    (render_sharded_peer; CUDA, device-resident rays), "gather" = NCCL all_gather_into_tensor of 16 B/ray + one
    permuted copy, "auto" = "peer" when torch symmetric memory is available, else "gather"."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = rays.shape[0]
    if gather and world > 1 and placement in ("auto", "peer") and rays.is_cuda and hasattr(tensorf, "render_eval") \
            and world <= 8 and not render_kw.get("is_train", False):
        try:
            return render_sharded_peer(rays, tensorf, group=group, tile=tile, device=device, **render_kw)
        except (ImportError, RuntimeError, AttributeError) as exc:
            if placement == "peer":
                raise
            global _peer_fallback_reason
            _peer_fallback_reason = f"{type(exc).__name__}: {exc}"
    mine, idx = local_rays(rays, world, rank, tile)
    kw = dict(render_kw)
    if device is not None:
        kw["device"] = device
    rgb, _, depth, _, _ = renderer(mine, tensorf, **kw)
    if not gather or world == 1:
        if world == 1:
            return rgb, depth
        return rgb, depth, idx
    # equal-sized payloads for all_gather_into_tensor: every rank pads its shard to max_tiles full tiles, so the
    # gathered buffer is [world][max_tiles][tile] and ONE permuted copy puts tile t = j * world + r back at offset
    # t * tile (the only partial tile is the last one of the image, so truncating to n is enough)
    n_tiles = (n + tile - 1) // tile
    max_tiles = (n_tiles + world - 1) // world
    per = max_tiles * tile
    payload = torch.zeros((per, 4), dtype=rgb.dtype, device=rgb.device)
    payload[: rgb.shape[0], :3] = rgb
    payload[: rgb.shape[0], 3] = depth
    full = torch.empty((world * per, 4), dtype=rgb.dtype, device=rgb.device)
    dist.all_gather_into_tensor(full, payload, group=group)
    out = full.view(world, max_tiles, tile, 4).permute(1, 0, 2, 3).reshape(-1, 4)[:n]
    return out[:, :3].contiguous(), out[:, 3].contiguous()


class GradSync:
    """All-reduce (SUM or MEAN) of the gradient buckets of a replicated TensorVMSplit.

    Attach with `GradSync(model, group).install()`: the packed factor-gradient buffer is reduced inside the
    autograd backward (before it is unpacked), and `finish()` — called after loss.backward() — reduces the
    small basis/MLP bucket.  With losses normalised by the LOCAL ray count use average=True (global mean)."""

    def __init__(self, model, group=None, average: bool = True, transport: str = "auto"):
        """transport: "peer" = the gradient workspace lives in NVLink peer memory (torch symmetric memory) and is reduced
        by the library's own two-shot kernel (tvm_allreduce_sum_peer: NVSwitch multicast ld_reduce / st when available,
        else peer loads and stores); "nccl" = dist.all_reduce; "auto" = "peer" when symmetric memory can be set up on a
        CUDA/NCCL group of <= 8 ranks, else "nccl"."""
        self.model, self.group, self.average = model, group, average
        self.transport = transport
        self.calls = 0
        self.bytes = 0
        self._small_done = False
        self._peer = None             # {"buf", "hdl", "ptrs", "mc"} once the peer workspace exists
        self.transport_used = "nccl"

    # ---- peer-memory workspace (called by TensorVMSplit.grad_workspace while this sync is installed)
    def alloc_workspace(self, n_floats: int, device):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        want_peer = self.transport in ("auto", "peer") and world > 1 and world <= 8 and torch.device(device).type == "cuda"
        if want_peer:
            try:
                import torch.distributed._symmetric_memory as symm
                pg = self.group if self.group is not None else dist.group.WORLD
                n = (n_floats + 3) // 4 * 4
                buf = symm.empty((n,), dtype=torch.float32, device=device)
                buf.zero_()
                hdl = symm.rendezvous(buf, pg.group_name)
                self._peer = {"buf": buf, "hdl": hdl, "ptrs": [int(p) for p in hdl.buffer_ptrs],
                              "mc": int(getattr(hdl, "multicast_ptr", 0) or 0)}
                hdl.barrier(channel=0)
                self.transport_used = "peer-multicast" if self._peer["mc"] else "peer"
                return buf[:n_floats]
            except (ImportError, RuntimeError, AttributeError) as exc:
                if self.transport == "peer":
                    raise
                self.transport_used = f"nccl ({type(exc).__name__}: {exc})"[:200]
        self._peer = None
        return torch.zeros(n_floats, dtype=torch.float32, device=device)

    def install(self):
        self.model.grad_sync = self
        return self

    def remove(self):
        self.model.grad_sync = None

    def _reduce(self, flat: torch.Tensor):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            flat.div_(dist.get_world_size(self.group))
        self.calls += 1
        self.bytes += flat.numel() * flat.element_size()
        return flat

    def reduce_sum(self, flat: torch.Tensor, covers_small_params: bool = False) -> float:
        """SUM all-reduce of a flat gradient workspace in place (no division); returns the factor the caller applies while
        unpacking (1/world when averaging).  covers_small_params: the workspace also holds the basis / MLP gradients of
        this step, so finish() has nothing left to reduce."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1:
            pk = self._peer
            if pk is not None and flat.data_ptr() == pk["buf"].data_ptr() and flat.numel() <= pk["buf"].numel():
                import ctypes as C
                from . import _lib
                lib = _lib.load()
                rank = dist.get_rank(self.group)
                pk["hdl"].barrier(channel=0)             # every rank's scatter into its workspace has completed
                with torch.cuda.device(flat.device):
                    _lib.check(lib.tvm_allreduce_sum_peer(_lib.ptr_array_int(pk["ptrs"]), world, rank, pk["buf"].numel(),
                                                          C.c_void_p(pk["mc"] or None),
                                                          C.c_void_p(torch.cuda.current_stream(flat.device).cuda_stream)),
                               "tvm_allreduce_sum_peer")
                pk["hdl"].barrier(channel=0)             # every slice has been written into every workspace
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            self.calls += 1
            self.bytes += flat.numel() * flat.element_size()
        self._small_done = covers_small_params
        return 1.0 / world if (self.average and world > 1) else 1.0

    def reduce_packed_factor_grads(self, g_packed: torch.Tensor):
        """Called by the backward with the flat packed factor-gradient buffer (in place)."""
        return self._reduce(g_packed)

    def small_params(self):
        m = self.model
        return [m.basis_mat.weight] + list(m.renderModule.parameters())

    def finish(self):
        """Reduce basis_mat + MLP gradients as one flat bucket (call once per step, after backward).  A no-op when the
        backward already reduced them together with the factor gradients (MLP_Fea head: one workspace, one all-reduce)."""
        if self._small_done:
            self._small_done = False
            return
        ps = [p for p in self.small_params() if p.grad is not None]
        if not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        self._reduce(flat)
        off = 0
        for p in ps:
            k = p.grad.numel()
            p.grad.copy_(flat[off:off + k].view_as(p.grad))
            off += k
