"""CUDA-graph capture of launch-bound steps (iffnerf_b200.graphs.CapturedStep): replays must reproduce eager steps."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _pose_setup(dev, seed):
    import iffnerf_b200 as I
    fld, _ = fx.config1(0.0, "sphere", 6)
    m = H.module_from_field(fld, dev)
    for p in m.parameters():
        p.requires_grad_(False)
    Hh = Ww = 100
    focal = 0.5 * Ww / np.tan(0.5 * 0.6911112)
    K = torch.tensor([[[focal, 0.0, Ww / 2], [0.0, focal, Hh / 2], [0.0, 0.0, 1.0]]], dtype=torch.float32)
    base = torch.cat([fx.orbit_pose(40.0, 25.0), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0).to(dev)
    delta = torch.zeros(3, 4, device=dev, requires_grad=True)
    opt = torch.optim.Adam([delta], lr=1e-3, capturable=True)
    gen = torch.Generator().manual_seed(seed)
    batches = [(torch.stack([torch.randint(25, 75, (256,), generator=gen), torch.randint(25, 75, (256,), generator=gen)], -1)
                .to(torch.int32), torch.rand(256, 3, generator=gen)) for _ in range(4)]
    static = {"pixels": batches[0][0].to(dev), "target": batches[0][1].to(dev), "bg": torch.tensor([0.3, 0.6, 0.9], device=dev)}

    def step():
        opt.zero_grad(set_to_none=True)
        pose = base + torch.cat([delta, torch.zeros(1, 4, device=dev)], 0)
        rays = I.pixel_rays(K, pose, static["pixels"])
        rgb = m(rays, bg_color=static["bg"], is_train=False)[0]
        loss = torch.mean((rgb - static["target"]) ** 2)
        loss.backward()
        opt.step()
        return loss
    return m, delta, static, batches, step


def test_captured_pose_refinement_step_matches_eager(built_lib):
    """An iNeRF-style iteration (fused ray generation -> render -> MSE -> backward to the pose -> Adam) replayed from a
    CUDA graph follows the same loss / parameter trajectory as the eager loop on the same batches (both start from the
    same two warm-up steps, which also initialise the optimiser state outside the capture)."""
    import iffnerf_b200 as I
    dev = torch.device("cuda:0")
    m, delta, static, batches, step = _pose_setup(dev, 5)
    for _ in range(2):
        step()
    eager_losses = []
    for pix, tgt in batches:
        static["pixels"].copy_(pix); static["target"].copy_(tgt)
        eager_losses.append(step().item())

    m2, delta2, static2, batches2, step2 = _pose_setup(dev, 5)
    graphed = I.graphs.CapturedStep(step2, models=[m2], warmup=2)
    losses = []
    for pix, tgt in batches2:
        static2["pixels"].copy_(pix); static2["target"].copy_(tgt)
        losses.append(graphed().item())
    np.testing.assert_allclose(losses, eager_losses, rtol=1e-5, atol=1e-7)
    assert delta2.detach().abs().max().item() > 0                 # the captured optimiser step moves the pose
    assert (delta2.detach() - delta.detach()).abs().max().item() <= 1e-6


def test_captured_train_step_matches_eager(built_lib):
    """train.py-style step (4096-ray shape reduced to 512) with Adam(capturable=True): three replays == three eager steps."""
    import iffnerf_b200 as I
    dev = torch.device("cuda:0")

    def setup():
        fld, rays = fx.config1(0.0, "sphere", 6)
        m = H.module_from_field(fld, dev)
        m.train()
        opt = torch.optim.Adam(m.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99), capturable=True)
        sub, _ = fx.subsample(rays, 512, seed=9)
        g = torch.Generator().manual_seed(1)
        static = {"rays": sub.to(dev), "target": torch.rand(512, 3, generator=g).to(dev),
                  "jitter": torch.rand(512, generator=g).to(dev), "bg": torch.ones(3, device=dev)}

        def step():
            opt.zero_grad(set_to_none=True)
            rgb, _, _, alpha, _, _ = m(static["rays"], bg_color=static["bg"], is_train=True, N_samples=443,
                                       jitter=static["jitter"])
            loss = torch.mean((rgb - static["target"]) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))
            loss.backward()
            opt.step()
            return loss
        return m, opt, step

    m, opt, step = setup()
    step()                                                              # same warm-up step as the captured run
    eager = [step().item() for _ in range(3)]
    after_eager = {k: v.detach().clone() for k, v in m.state_dict().items()}

    m2, opt2, step2 = setup()
    graphed = I.graphs.CapturedStep(step2, models=[m2], warmup=1)      # warm-up initialises Adam's state eagerly
    replay = [graphed().item() for _ in range(3)]
    np.testing.assert_allclose(replay, eager, rtol=2e-5, atol=1e-7)
    for k, v in m2.state_dict().items():
        ref = after_eager[k]
        assert (v - ref).abs().max().item() <= 1e-5 + 2e-4 * ref.abs().max().item(), k
    # the shadows were invalidated: an eager eval render after the replays uses the UPDATED parameters
    _, rays = fx.config1(0.0, "sphere", 6)
    with torch.no_grad():
        a = m2.render_eval(rays[:2048].to(dev), white_bg=True)["rgb_map"]
        b = m.render_eval(rays[:2048].to(dev), white_bg=True)["rgb_map"]
    assert (a - b).abs().max().item() <= 2e-4
