"""iffnerf_b200/synthetic.py (the workloads bench.py renders, product classes only) == oracle/fixtures.py (what the
goldens were generated from): same rays bit for bit, same parameters for the same seed, same occupancy volume."""
import torch

from iffnerf_b200 import synthetic as syn
from oracle import fixtures as fx


def test_rays_and_poses_identical():
    assert torch.equal(syn.config2_rays(40, 56, 80.0, 20.0), fx.config2_rays(40, 56, 80.0, 20.0))
    assert torch.equal(syn.orbit_pose(125.0), fx.orbit_pose(125.0))
    c2w = syn.look_at_c2w((2.2, 1.6, 0.9), target=(0.0, 0.0, 0.25))
    assert torch.equal(syn.pinhole_rays(9, 16, 14.4, c2w, cols=6), fx.pinhole_rays(9, 16, 14.4, c2w, cols=6))


def test_model_parameters_and_occupancy_identical():
    grid = [20, 24, 18]
    aabb = torch.tensor(syn.TRUCK_AABB)
    fld = fx.make_field(grid, aabb=aabb, near_far=(0.01, 6.0), occ_res=(12, 14, 10))
    m = syn.build_model(grid, "cpu", aabb=aabb, near_far=(0.01, 6.0), occ_res=(12, 14, 10))
    sd = m.state_dict()
    for k in range(3):
        assert torch.equal(sd[f"density_plane.{k}"], fld.density_plane[k])
        assert torch.equal(sd[f"density_line.{k}"], fld.density_line[k])
        assert torch.equal(sd[f"app_plane.{k}"], fld.app_plane[k])
        assert torch.equal(sd[f"app_line.{k}"], fld.app_line[k])
    assert torch.equal(sd["basis_mat.weight"], fld.basis)
    for i, li in enumerate((0, 2, 4)):
        assert torch.equal(sd[f"renderModule.mlp.{li}.weight"], fld.mlp_w[i])
        assert torch.equal(sd[f"renderModule.mlp.{li}.bias"], fld.mlp_b[i])
    assert torch.equal(m.alphaMask.alpha_volume.reshape(fld.occupancy.volume.shape), fld.occupancy.volume)
    assert syn.n_to_reso(300 ** 3, aabb) == fx.config4(2, 2)[0].grid
