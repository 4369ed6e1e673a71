"""Seeded synthetic workloads of the BASELINE.json configs built from the PRODUCT classes only — what
bench.py, the scripts and the examples render.  There is no dataset or checkpoint on the box, so the workloads are
random-init fields of the reference's shapes plus analytic cameras.

`tests/test_synthetic_fixtures.py` pins these builders to the oracle's own fixture builders (oracle/fixtures.py): same
rays bit for bit, same parameters for the same seed (the constructor draws from torch's RNG in the reference's order).
"""
from __future__ import annotations

import contextlib
import io
import math

import torch

SEED = 20211202          # train.py:509
TRUCK_AABB = [[-1.35, -1.10, -0.55], [1.32, 1.14, 1.12]]
FOV_X = 0.6911112        # Blender lego camera_angle_x


def look_at_c2w(cam_pos, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """OpenCV-style camera-to-world [3,4] (x right, y down, z forward)."""
    p = torch.tensor(cam_pos, dtype=torch.float32)
    fwd = torch.tensor(target, dtype=torch.float32) - p
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, torch.tensor(up, dtype=torch.float32))
    if right.norm() < 1e-6:
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    return torch.stack([right, down, fwd, p], dim=1)


def orbit_pose(theta_deg=35.0, phi_deg=30.0, radius=4.03):
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    return look_at_c2w((radius * math.cos(ph) * math.cos(th), radius * math.cos(ph) * math.sin(th),
                        radius * math.sin(ph)))


def pinhole_rays(H, W, focal, c2w, cols=7, cx=None, cy=None):
    """Rays in the layout the reference loaders emit: [H*W, 7] = (o, unit d, radii) (dataLoader/blender.py:105-114,
    radii as in ray_utils.py:90-98) or [H*W, 6]."""
    cx = W / 2 if cx is None else cx
    cy = H / 2 if cy is None else cy
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5, torch.arange(W, dtype=torch.float32) + 0.5,
                          indexing="ij")

    def cam_dir(ii, jj):
        return torch.stack([(ii - cx) / focal, (jj - cy) / focal, torch.ones_like(ii)], -1)

    rot = c2w[:3, :3]
    d0, dx, dy = cam_dir(i, j) @ rot.T, cam_dir(i + 1, j) @ rot.T, cam_dir(i, j + 1) @ rot.T
    radii = 0.5 * ((dx - d0).norm(dim=-1) + (dy - d0).norm(dim=-1)) * (2 / math.sqrt(12))
    d = d0 / d0.norm(dim=-1, keepdim=True)
    parts = [c2w[:3, 3].expand_as(d).reshape(-1, 3), d.reshape(-1, 3)]
    if cols == 7:
        parts.append(radii.reshape(-1, 1))
    return torch.cat(parts, -1).contiguous()


def sphere_volume(aabb, res=200, radius=1.0):
    """{0,1} occupancy volume [Dz][Dy][Dx] on a lattice spanning `aabb`: 1 inside |x| < radius."""
    res3 = (res, res, res) if isinstance(res, int) else tuple(res)
    xs = [torch.linspace(float(aabb[0][a]), float(aabb[1][a]), res3[a]) for a in range(3)]
    zz, yy, xx = torch.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
    return ((xx * xx + yy * yy + zz * zz) < radius * radius).float().contiguous()


def n_to_reso(n_voxels, aabb):
    """utils.py:20-24: grid with ~n_voxels cells of equal edge length inside the box."""
    lo, hi = aabb
    vox = ((hi - lo).prod() / n_voxels).pow(1 / len(lo))
    return ((hi - lo) / vox).long().tolist()


def build_model(grid, device, aabb=None, density_shift=0.0, near_far=(2.0, 6.0), occ_res=200, shading="MLP_Fea",
                seed=SEED, view_pe=2, fea_pe=2):
    """Lego-shaped TensorVMSplit (16x3 density / 48x3 appearance components, app_dim 27) with random-init parameters
    (one seed -> the reference constructor's parameters) and a sphere occupancy of radius min(box)/3."""
    from .tensorf import AlphaGridMask, TensorVMSplit
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3]) if aabb is None else aabb
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = TensorVMSplit(aabb.clone().to(device), [int(g) for g in grid], device, density_n_comp=[16] * 3,
                          appearance_n_comp=[48] * 3, app_dim=27, near_far=list(near_far), shadingMode=shading,
                          alphaMask_thres=1e-4, density_shift=density_shift, distance_scale=25, pos_pe=6,
                          view_pe=view_pe, fea_pe=fea_pe, featureC=128, step_ratio=0.5, fea2denseAct="softplus")
    if occ_res is not None:
        radius = float((aabb[1] - aabb[0]).min()) / 3
        m.alphaMask = AlphaGridMask(device, aabb.clone().to(device), sphere_volume(aabb, occ_res, radius).to(device))
    return m


def config2_rays(H=800, W=800, theta_deg=35.0, phi_deg=30.0):
    """BASELINE configs[1]: Blender-style 7-column rays of one orbit view."""
    return pinhole_rays(H, W, 0.5 * W / math.tan(0.5 * FOV_X), orbit_pose(theta_deg, phi_deg), cols=7)


def config2_model(device, density_shift=0.0):
    return build_model([300] * 3, device, density_shift=density_shift)


def config4(device, H=1080, W=1920):
    """BASELINE configs[3]: non-cubic Tanks&Temples-like box, ~300^3 voxels, 1920x1080 rays."""
    aabb = torch.tensor(TRUCK_AABB)
    m = build_model(n_to_reso(300 ** 3, aabb), device, aabb=aabb, near_far=(0.01, 6.0), occ_res=(180, 200, 160))
    rays = pinhole_rays(H, W, 0.9 * W, look_at_c2w((2.2, 1.6, 0.9), target=(0.0, 0.0, 0.25)), cols=7)
    return m, rays
