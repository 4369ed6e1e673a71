"""Host-side logic of the multi-GPU path on CPU: cyclic tile sharding, result reassembly and the gradient
buckets, exercised with world_size=2 over the gloo backend (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from iffnerf_b200 import sharding


def test_shard_tiles_partition_every_ray_exactly_once():
    for n, world, tile in ((640000, 8, 4096), (10000, 2, 4096), (4095, 4, 4096), (1, 2, 16), (0, 2, 16),
                           (2073600, 8, 4096)):
        seen = torch.zeros(n, dtype=torch.int32)
        sizes = []
        for r in range(world):
            idx = sharding.shard_index(n, world, r, tile)
            seen[idx] += 1
            sizes.append(idx.numel())
            tiles = sharding.shard_tiles(n, world, r, tile)
            assert idx.numel() == sum(b - a for a, b in tiles)
            ref = torch.cat([torch.arange(a, b) for a, b in tiles]) if tiles else torch.empty(0, dtype=torch.long)
            assert torch.equal(idx, ref)                    # the vectorised index == the concatenated tile ranges
        assert bool((seen == 1).all())
        assert max(sizes) - min(sizes) <= tile              # balanced to within one tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeField(torch.nn.Module):
    """Stands in for TensorVMSplit on CPU: 'renders' a deterministic function of each ray."""
    def __init__(self):
        super().__init__()
        self.basis_mat = torch.nn.Linear(4, 3, bias=False)
        self.renderModule = torch.nn.Linear(3, 3)
        self.grad_sync = None


def _fake_renderer(rays, tensorf, **kw):
    rgb = torch.stack([rays[:, 0], rays[:, 1] * 2, rays[:, 0] + rays[:, 1]], -1)
    return rgb, None, rays[:, 2] * 3, None, None


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        rays = torch.rand(1000, 6)
        # ---- forward: every rank gets the full image, in the original ray order
        rgb, depth = sharding.render_sharded(rays, _FakeField(), _fake_renderer, tile=64, gather=True)
        ref_rgb, _, ref_depth, _, _ = _fake_renderer(rays, None)
        assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
        # local-only mode returns the cyclic slice
        lrgb, ldepth, idx = sharding.render_sharded(rays, _FakeField(), _fake_renderer, tile=64, gather=False)
        assert torch.equal(lrgb, ref_rgb[idx]) and idx.numel() == len(sharding.shard_index(1000, world, rank, 64))
        # ---- training: packed factor bucket + small bucket are averaged across ranks
        m = _FakeField()
        sync = sharding.GradSync(m, average=True).install()
        packed = torch.full((1000,), float(rank + 1))
        sync.reduce_packed_factor_grads(packed)
        assert torch.allclose(packed, torch.full((1000,), (1 + world) / 2))
        for i, p in enumerate(sync.small_params()):
            p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
        sync.finish()
        for i, p in enumerate(sync.small_params()):
            assert torch.allclose(p.grad, torch.full_like(p, (i + 1) * (1 + world) / 2))
        assert sync.calls == 2 and sync.bytes == 1000 * 4 + sum(p.numel() for p in sync.small_params()) * 4
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_render_and_grad_buckets(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
