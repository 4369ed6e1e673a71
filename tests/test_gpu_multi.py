"""Multi-GPU path on real devices (needs >= 2 GPUs; skipped otherwise): ray-sharded rendering over NCCL reproduces
the single-GPU image exactly, and a data-parallel train step with the packed-gradient all-reduce reproduces the
single-GPU gradients of the full batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    import iffnerf_b200 as I
    from iffnerf_b200 import sharding
    from oracle import fixtures as fx
    from tests import helpers as H
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        fld, rays = fx.config1(0.0, "sphere", 7)
        m = H.module_from_field(fld, dev)
        # ---- sharded eval: every rank ends up with the full image, identical to a single-GPU render
        ref_rgb, _, ref_depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev)
        rgb, depth = sharding.render_sharded(rays, m, I.OctreeRender_trilinear_fast, tile=512, gather=True,
                                             device=dev, white_bg=True, placement="gather")
        assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
        # ---- the same with the shading epilogue storing into every rank's image over NVLink peer memory; three frames
        # so both buffers of the double-buffered pair are reused once
        rays_dev = rays.to(dev)
        for _ in range(3):
            rgb, depth = sharding.render_sharded(rays_dev, m, I.OctreeRender_trilinear_fast, tile=512, gather=True,
                                                 device=dev, white_bg=True, placement="peer")
            assert torch.equal(rgb, ref_rgb) and torch.equal(depth, ref_depth)
        # ---- data-parallel train step == single-GPU step on the whole batch
        torch.manual_seed(3)
        batch = rays[torch.randperm(rays.shape[0])[:2048]].to(dev)
        jit = torch.rand(2048, device=dev)
        target = torch.rand(2048, 3, device=dev)

        def step(sel, sync):
            m.zero_grad()
            rgbm, _, _, alpha, _, _ = m(batch[sel], bg_color=torch.ones(3, device=dev), is_train=True, jitter=jit[sel])
            loss = torch.mean((rgbm - target[sel]) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))
            loss.backward()
            if sync is not None:
                sync.finish()
            return [p.grad.clone() for p in m.parameters()]

        full = step(torch.arange(2048, device=dev), None)
        mine = sharding.shard_index(2048, world, rank, tile=2048 // world, device=dev)
        # "peer": the library's own all-reduce kernel over NVLink peer memory; "nccl": dist.all_reduce
        for transport in ("peer", "nccl"):
            sync = sharding.GradSync(m, average=True, transport=transport).install()
            for _ in range(2):          # twice: the workspace must come back zeroed from the first step
                sharded = step(mine, sync)
            sync.remove()
            assert sync.calls == 2      # ONE all-reduce per step: factors + MLP + basis share the gradient workspace
            assert sync.transport_used.startswith(transport)
            for a, b in zip(sharded, full):
                scale = b.abs().max().item()
                assert (a - b).abs().max().item() <= 1e-4 * scale + 1e-9, (transport, (a - b).abs().max().item(), scale)
        torch.cuda.synchronize()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_sharded_render_and_data_parallel_step(tmp_path, built_lib):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
