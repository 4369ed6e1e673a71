"""SURVEY.md §8f "next" rows on the GPU: point / short-ray queries, occupancy-grid maintenance, resizing."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc
from oracle.make_golden import point_rays
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def c1(dev):
    fld, rays = fx.config1(0.0, "sphere", 6)
    return fld, rays, H.module_from_field(fld, dev)


def test_sample_point_color_forward_matches_reference_golden(c1, dev):
    """pose_estimation/sampling.py:245-250: model(rays6, N_samples=20, sample_func=model.sample_point_color)."""
    fld, _, m = c1
    g = H.golden("c1_point20")
    H.check_params(fld, g)
    rays6 = point_rays(fld, 4096)
    with torch.no_grad():
        rgb, depth, acc, alpha, z, dists = m(rays6.to(dev), N_samples=20, sample_func=m.sample_point_color,
                                             white_bg=True)
    assert z.shape == (1, 20) and dists.shape == (1, 20)
    assert np.array_equal(z.cpu().numpy(), g["z_vals"]) and np.array_equal(dists.cpu().numpy(), g["dists"])
    assert np.abs(alpha.cpu().numpy() - g["alpha"]).max() <= 1e-5
    assert np.array_equal(alpha.cpu().numpy() > 0, g["alpha"] > 0)            # same valid samples
    assert np.abs(rgb.cpu().numpy() - g["rgb_map"]).max() <= 1e-4
    assert np.abs(depth.cpu().numpy() - g["depth_map"]).max() <= 1e-4
    assert np.abs(acc.cpu().numpy() - g["acc_map"]).max() <= 1e-4
    # the host-side sampler itself returns the reference's tensors
    pts, step, valid = m.sample_point_color(rays6[:8, :3].to(dev), rays6[:8, 3:6].to(dev), None, N_samples=20)
    ref = orc.sample_around_points(fld, rays6[:8, :3], rays6[:8, 3:6], 20)
    assert torch.equal(pts.cpu(), ref[0]) and torch.equal(step.cpu(), ref[1]) and torch.equal(valid.cpu(), ref[2])
    # other callables are refused, not silently mis-rendered
    with pytest.raises(NotImplementedError), torch.no_grad():
        m(rays6[:4].to(dev), sample_func=lambda *a, **k: None)


def test_point_queries_match_reference_golden(c1, dev):
    fld, _, m = c1
    g = H.golden("c1_point20")
    pts = torch.from_numpy(g["points"]).to(dev)
    a = m.compute_alpha(pts, length=m.stepSize.item())
    assert np.abs(a.cpu().numpy() - g["point_alpha"]).max() <= 1e-6
    assert np.array_equal(a.cpu().numpy() > 0, g["point_alpha"] > 0)          # occupancy gate is exact
    f = m.compute_densityfeature(m.normalize_coord(pts))
    assert np.abs(f.cpu().numpy() - g["point_feature"]).max() <= 2e-5
    # points outside the box: zero padding, like F.grid_sample
    far = torch.tensor([[3.0, 0.0, 0.0], [0.0, -2.5, 0.1], [1.4999, 1.4999, -1.4999]], device=dev)
    ref = orc.density_feature(fld, orc.normalize(fld, far.cpu()))
    got = m.compute_densityfeature(m.normalize_coord(far))
    assert (got.cpu() - ref).abs().max() <= 2e-5
    # compute_appfeature (tensoRF.py:237-256; pose_estimation/sampling.py:535-541): [M, app_dim] incl. basis_mat
    af = m.compute_appfeature(m.normalize_coord(pts))
    assert af.shape == (pts.shape[0], 27)
    assert np.abs(af.cpu().numpy() - g["point_appfeature"]).max() <= 2e-5
    ref = orc.app_feature(fld, orc.normalize(fld, far.cpu()))
    assert (m.compute_appfeature(m.normalize_coord(far)).cpu() - ref).abs().max() <= 2e-5


def test_update_alpha_mask_and_filtering_match_oracle(dev):
    fld, rays = fx.config1(0.0, None, 6)
    m = H.module_from_field(fld, dev)
    grid = (48, 40, 56)
    new_aabb = m.updateAlphaMask(grid)
    vol_ref, aabb_ref = orc.dense_alpha_volume(fld, grid, thres=1e-4)
    vol = m.alphaMask.alpha_volume.view(vol_ref.shape).cpu()
    assert vol.shape == (56, 40, 48)
    assert (vol != vol_ref).float().mean().item() <= 1e-3       # threshold-borderline voxels only
    assert (new_aabb.cpu() - aabb_ref).abs().max() <= 0.1
    # ray filtering against the new occupancy volume == the reference's sample-then-test
    fld.occupancy = orc.OccupancyGrid(aabb=fld.aabb.clone(), volume=vol.clone())
    sub = rays[::5].contiguous()
    rgbs = torch.zeros(sub.shape[0], 3)
    kept, _ = m.filtering_rays(sub, rgbs, N_samples=256)
    pts, _, valid = orc.sample_along_rays(fld, sub[:, :3], sub[:, 3:6], 256)
    occ = (orc.occupancy_value(fld.occupancy, pts.view(-1, 3)) > 0).view(valid.shape)
    assert torch.equal(kept, sub[(occ & valid).any(-1)])
    kept_b, _ = m.filtering_rays(sub, rgbs, bbox_only=True)
    assert kept_b.shape[0] >= kept.shape[0]


def _field_from_module(m, fld):
    import copy
    f2 = copy.copy(fld)
    f2.aabb = m.aabb.detach().cpu().clone()
    f2.grid = m.gridSize.tolist()
    f2.density_plane = [p.detach().cpu() for p in m.density_plane]
    f2.density_line = [p.detach().cpu() for p in m.density_line]
    f2.app_plane = [p.detach().cpu() for p in m.app_plane]
    f2.app_line = [p.detach().cpu() for p in m.app_line]
    return f2


def test_upsample_and_shrink_keep_rendering_in_parity(dev):
    """train.py:384-415: upsample_volume_grid / shrink replace the Parameters; the packed shadow follows."""
    fld = fx.make_field([48, 48, 48], density_shift=0.0, holes_seed=1)
    m = H.module_from_field(fld, dev)
    rays = fx.config1(0.0, None, 6)[1][3000:3600].contiguous()
    _ = m.render_eval(rays.to(dev), white_bg=True)                      # populate the caches first
    m.upsample_volume_grid([64, 72, 80])
    assert tuple(m.density_plane[0].shape) == (1, 16, 72, 64) and m.gridSize.tolist() == [64, 72, 80]
    f_up = _field_from_module(m, fld)
    got = m.render_eval(rays.to(dev), white_bg=True, early_term=False)
    with torch.no_grad():
        ref = orc.render_chunk(f_up, rays, white_bg=True)
    assert m.nSamples == orc.step_geometry(f_up.aabb, f_up.grid, 0.5)["nSamples"]
    assert (got["rgb_map"].cpu() - ref["rgb_map"]).abs().max() <= 1e-4
    m.shrink(torch.tensor([[-1.0, -0.9, -1.1], [1.1, 1.0, 0.9]]))
    f_sh = _field_from_module(m, fld)
    got = m.render_eval(rays.to(dev), white_bg=True, early_term=False)
    bits, _ = m.sample_mask(rays.to(dev))
    with torch.no_grad():
        ref = orc.render_chunk(f_sh, rays, white_bg=True)
    assert np.array_equal(H.unpack_bits(bits.cpu().numpy(), m.nSamples), ref["ray_valid"].numpy())
    assert (got["rgb_map"].cpu() - ref["rgb_map"]).abs().max() <= 1e-4


def test_regularisers_add_into_the_same_grads(dev):
    """train.py:299-325: torch regularisers on the raw factors accumulate with the kernels' gradients."""
    fld = fx.make_field([32, 32, 32], density_shift=0.0, occupancy=None)
    m = H.module_from_field(fld, dev)
    rays = fx.config1(0.0, None, 6)[1][4000:4128].to(dev)
    m.zero_grad()
    reg = m.vector_comp_diffs() * 1e-3 + m.density_L1() * 1e-4
    reg.backward()
    g_reg = m.density_plane[0].grad.clone()
    rgb = m(rays, white_bg=True, is_train=True)[0]
    rgb.mean().backward()
    assert torch.isfinite(m.density_plane[0].grad).all()
    assert not torch.equal(m.density_plane[0].grad, g_reg)             # render gradient was added on top
