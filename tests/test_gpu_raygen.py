"""SURVEY.md §8f row 4 on the GPU: fused pixel -> ray generation with the pose gradient, through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_pixel_rays_match_reference_golden(dev):
    """Forward within fp32 rounding of the reference's get_ray_directions_Ks + get_rays + F.normalize chain
    (inerf/estimate_pose_inerf.py:96-99,149-164); pose gradient of a fixed linear functional within 1e-5 relative."""
    import iffnerf_b200 as I
    g = H.golden("c5_raygen")
    pose = torch.from_numpy(g["c2w"]).to(dev).requires_grad_(True)
    rays = I.pixel_rays(torch.from_numpy(g["K"]), pose, torch.from_numpy(g["pixels"]))
    assert rays.shape == (1024, 7)
    (rays * torch.from_numpy(g["upstream"]).to(dev)).sum().backward()
    r = rays.detach().cpu().numpy()
    assert np.abs(r[:, :3] - g["rays"][:, :3]).max() == 0.0                      # origins are copies
    assert np.abs(r[:, 3:6] - g["rays"][:, 3:6]).max() <= 2e-7                   # unit directions: 1-2 ulp
    # radii = |R dx - R ori|: a difference of nearly equal unit-scale vectors, so 1 ulp on them is ~1e-4 relative here
    assert np.abs(r[:, 6] - g["rays"][:, 6]).max() <= 3e-4 * np.abs(g["rays"][:, 6]).max()
    scale = np.abs(g["d_c2w"]).max()
    assert pose.grad.shape == (4, 4)
    assert np.abs(pose.grad.cpu().numpy() - g["d_c2w"]).max() <= 1e-5 * scale
    loader = I.pixel_rays(torch.from_numpy(g["K"]), torch.from_numpy(g["c2w"]).to(dev), torch.from_numpy(g["pixels"]),
                          renormalize=False).cpu().numpy()
    assert np.abs(loader[:, 3:6] - g["loader_rays"][:, 3:6]).max() <= 2e-7


def test_full_image_generation_matches_fixture_rays(dev):
    """pixels=None: every pixel of an HxW image in the loaders' row-major order (dataLoader/blender.py:105-114)."""
    import iffnerf_b200 as I
    g = H.golden("c5_raygen")
    Hh, Ww = 60, 80
    K = torch.tensor([[[90.0, 0.0, Ww / 2], [0.0, 90.0, Hh / 2], [0.0, 0.0, 1.0]]])
    c2w = torch.from_numpy(g["c2w"])
    ys, xs = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)
    ref = orc.pixel_rays(K, c2w, pix, Hh, Ww, renormalize=False)
    out = I.pixel_rays(K, c2w.to(dev), None, image_wh=(Ww, Hh), renormalize=False).cpu()
    assert out.shape == (Hh * Ww, 7)
    assert (out - ref).abs().max() <= 1e-6


def test_batched_candidate_poses_chain_to_pose_parameters(dev):
    """BASELINE config 5: several candidate poses in ONE call; gradients of a render loss flow through the fused ray
    generation into every pose matrix and agree with oracle autograd through its own ray build + renderer."""
    import iffnerf_b200 as I
    fld, _ = fx.config1(0.0, "sphere", 6)
    m = H.module_from_field(fld, dev)
    for p in m.parameters():
        p.requires_grad_(False)
    Hh = Ww = 100
    focal = 0.5 * Ww / np.tan(0.5 * 0.6911112)
    K = torch.tensor([[[focal, 0.0, Ww / 2], [0.0, focal, Hh / 2], [0.0, 0.0, 1.0]]], dtype=torch.float32)
    gen = torch.Generator().manual_seed(3)
    P, per = 3, 96
    poses = torch.stack([torch.cat([fx.orbit_pose(30.0 + 25 * k, 25.0), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
                         for k in range(P)])
    pix = torch.stack([torch.randint(30, 70, (P * per,), generator=gen), torch.randint(30, 70, (P * per,), generator=gen)], -1)
    pidx = torch.arange(P).repeat_interleave(per)
    target = torch.rand(P * per, 3, generator=gen)
    bg = torch.rand(3, generator=gen)

    pd = poses.to(dev).requires_grad_(True)
    rays = I.pixel_rays(K, pd, pix, pose_index=pidx)
    rgb = m(rays, bg_color=bg.to(dev), is_train=False)[0]
    torch.mean((rgb - target.to(dev)) ** 2).backward()

    po = poses.clone().requires_grad_(True)
    rays_o = torch.cat([orc.pixel_rays(K, po[k], pix[pidx == k], Hh, Ww) for k in range(P)], 0)
    out = orc.render_chunk(fld, rays_o, bg_color=bg)
    torch.mean((out["rgb_map"] - target) ** 2).backward()
    assert (rays.detach().cpu() - rays_o.detach()).abs().max() <= 1e-6
    assert (rgb.detach().cpu() - out["rgb_map"].detach()).abs().max() <= 1e-4
    scale = po.grad.abs().max()
    assert scale > 0
    assert (pd.grad.cpu() - po.grad).abs().max() <= 5e-3 * scale


def test_pixel_rays_lie_equals_matrix_pose_and_backpropagates(dev):
    """`get_rays_lie` (ray_utils.py:103-140) takes the pose as a Lie-group element (kornia Se3: `.rotation.matrix()`,
    `.t`); pixel_rays_lie assembles [R | t] and must equal pixel_rays on the same matrix, with gradients reaching the
    group parameters (here an axis-angle vector through the matrix exponential)."""
    import types
    import iffnerf_b200 as I
    g = H.golden("c5_raygen")
    K, pix = torch.from_numpy(g["K"]), torch.from_numpy(g["pixels"])
    w = torch.tensor([0.02, -0.01, 0.03], device=dev, requires_grad=True)
    t = torch.from_numpy(g["c2w"][:3, 3]).to(dev).requires_grad_(True)
    base = torch.from_numpy(g["c2w"][:3, :3]).to(dev)

    def rotation():
        z = torch.zeros((), device=dev)
        skew = torch.stack([torch.stack([z, -w[2], w[1]]), torch.stack([w[2], z, -w[0]]), torch.stack([-w[1], w[0], z])])
        return torch.linalg.matrix_exp(skew) @ base
    se3 = types.SimpleNamespace(rotation=types.SimpleNamespace(matrix=rotation), t=t)
    rays = I.pixel_rays_lie(K, se3, pix)
    ref = I.pixel_rays(K, torch.cat([rotation(), t[:, None]], -1).detach(), pix)
    assert torch.equal(rays.detach(), ref)
    rays.square().sum().backward()
    assert w.grad is not None and t.grad is not None and float(w.grad.abs().max()) > 0 and float(t.grad.abs().max()) > 0
