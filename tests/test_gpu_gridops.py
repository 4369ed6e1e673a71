"""Grid maintenance and the `.th` checkpoint against goldens produced by the UNMODIFIED
reference (oracle/make_golden.py::gridops_case): updateAlphaMask -> filtering_rays -> shrink -> upsample_volume_grid,
then the reference-written checkpoint loaded by the product.  Masks, boxes, grid sizes and sample counts are exact;
resampled factors and renders carry fp32 tolerances written at the assertion."""
import os

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LATTICE = (48, 40, 56)
UPSAMPLE = [44, 38, 50]


def _fixture():
    aabb = torch.tensor(fx.TRUCK_AABB)
    fld = fx.make_field([32, 28, 36], aabb=aabb, near_far=(0.01, 6.0), occ_res=(30, 34, 26))
    c2w = fx.look_at_c2w((2.2, 1.6, 0.9), target=(0.0, 0.0, 0.25))
    return fld, fx.pinhole_rays(24, 40, 0.5 * 40, c2w, cols=7)


def _check_state(m, g, prefix, rtol):
    for k, v in m.state_dict().items():
        ref = g[f"{prefix}/{k}"]
        got = np.array([v.double().sum().item(), v.double().abs().sum().item()])
        assert abs(got[1] - ref[1]) <= rtol * ref[1], (k, got, ref)
        assert abs(got[0] - ref[0]) <= rtol * ref[1], (k, got, ref)


def _check_render(m, rays, g, prefix, dev):
    import iffnerf_b200 as I
    rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True, device=dev)
    _, counts = m.sample_mask(rays.to(dev), want_bits=False)
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy(), g[f"{prefix}_valid_count"])
    assert np.abs(rgb.cpu().numpy() - g[f"{prefix}_rgb"]).max() <= TOL
    assert np.abs(depth.cpu().numpy() - g[f"{prefix}_depth"]).max() <= TOL


def test_grid_maintenance_sequence_matches_reference(built_lib):
    dev = torch.device("cuda:0")
    g = H.golden("c_gridops")
    fld, rays = _fixture()
    H.check_params(fld, g)
    m = H.module_from_field(fld, dev)
    m.alphaMask_thres = float(g["thres"])
    # ---- updateAlphaMask (tensorBase.py:667-696): mask bits and the tight box are exact
    new_aabb = m.updateAlphaMask(LATTICE)
    vol = m.alphaMask.alpha_volume.reshape(LATTICE[::-1])
    assert tuple(vol.shape) == tuple(g["mask_shape"])
    assert np.array_equal(np.packbits(vol.bool().cpu().numpy().reshape(-1)), g["mask_bits"])
    assert np.array_equal(new_aabb.cpu().numpy(), g["mask_aabb"])
    # ---- filtering_rays (tensorBase.py:698-748), both modes: the kept-ray masks are exact
    index = torch.arange(rays.shape[0])[:, None].float()
    for mode, bbox_only in (("box", True), ("occ", False)):
        _, kept = m.filtering_rays(rays, index, N_samples=64, bbox_only=bbox_only)
        keep = np.zeros(rays.shape[0], dtype=bool)
        keep[kept.reshape(-1).long().numpy()] = True
        assert np.array_equal(keep, g[f"filter_{mode}"]), mode
    # ---- shrink (tensoRF.py:280-316): voxel range, snapped box and geometry scalars exact; factors are a crop
    m.shrink(new_aabb)
    assert m.gridSize.tolist() == g["shrink_grid"].tolist()
    assert np.array_equal(m.aabb.cpu().numpy(), g["shrink_aabb"])
    assert np.float32(m.stepSize.item()) == g["shrink_step"] and m.nSamples == int(g["shrink_nsamples"])
    _check_state(m, g, "shrink_sum", rtol=1e-12)
    _check_render(m, rays, g, "shrink", dev)
    # ---- upsample_volume_grid (tensoRF.py:258-278): bilinear align_corners=True, entries within 1e-6
    m.upsample_volume_grid(UPSAMPLE)
    assert m.gridSize.tolist() == g["up_grid"].tolist()
    assert np.float32(m.stepSize.item()) == g["up_step"] and m.nSamples == int(g["up_nsamples"])
    for k in range(3):
        for nme, p in ((f"density_plane.{k}", m.density_plane[k]), (f"app_line.{k}", m.app_line[k])):
            got = p.data.reshape(-1)[torch.from_numpy(g[f"up_idx/{nme}"]).to(dev)].cpu().numpy()
            assert np.abs(got - g[f"up_val/{nme}"]).max() <= 1e-6, nme
    _check_state(m, g, "up_sum", rtol=1e-6)
    _check_render(m, rays, g, "up", dev)


def test_reference_written_checkpoint_loads_and_renders(built_lib):
    """A `.th` written by the reference's save() (tensorBase.py:424-442) -> the product's constructor kwargs + load()
    -> the reference's own render of that model."""
    import iffnerf_b200 as I
    dev = torch.device("cuda:0")
    g = H.golden("c_gridops")
    _, rays = _fixture()
    ckpt = torch.load(os.path.join(GOLDEN, "c_ckpt_reference.th"), map_location=dev, weights_only=False)
    assert ckpt["model_name"] == "TensorVMSplit"
    kwargs = dict(ckpt["kwargs"])
    kwargs.update(device=dev)
    m = getattr(I, ckpt["model_name"])(**kwargs)
    m.load(ckpt)
    assert m.gridSize.tolist() == g["up_grid"].tolist()
    assert tuple(m.alphaMask.alpha_volume.shape[-3:]) == tuple(g["mask_shape"])
    assert np.array_equal(np.packbits(m.alphaMask.alpha_volume.bool().cpu().numpy().reshape(-1)), g["mask_bits"])
    _check_state(m, g, "up_sum", rtol=1e-12)
    _check_render(m, rays, g, "up", dev)
