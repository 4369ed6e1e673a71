"""TEST INFRASTRUCTURE ONLY — seeded synthetic fixtures for the BASELINE.json configs.

Pure torch/numpy; does not touch /root/reference, so it runs on the GPU box too.
"""
from __future__ import annotations

import math

import torch

from . import tensorf_oracle as orc

SEED = 20211202  # train.py:509

MODEL_KW = dict(n_sigma=(16, 16, 16), n_app=(48, 48, 48), app_dim=27, feature_c=128, view_pe=2, fea_pe=2)


def look_at_c2w(cam_pos, target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):
    """OpenCV-style camera-to-world [3,4] (x right, y down, z forward)."""
    p = torch.tensor(cam_pos, dtype=torch.float32)
    fwd = torch.tensor(target, dtype=torch.float32) - p
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float32)
    right = torch.linalg.cross(fwd, upv)
    if right.norm() < 1e-6:
        right = torch.linalg.cross(fwd, torch.tensor([0.0, 1.0, 0.0]))
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    return torch.stack([right, down, fwd, p], dim=1)


def orbit_pose(theta_deg=35.0, phi_deg=30.0, radius=4.03):
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    pos = (radius * math.cos(ph) * math.cos(th), radius * math.cos(ph) * math.sin(th), radius * math.sin(ph))
    return look_at_c2w(pos)


def pinhole_rays(H, W, focal, c2w, cols=7, cx=None, cy=None):
    """Rays in the layout the reference loaders emit: [H*W, 7] = (o, unit d, radii)
    (dataLoader/blender.py:105-114, radii as in ray_utils.py:90-98) or [H*W, 6]."""
    cx = W / 2 if cx is None else cx
    cy = H / 2 if cy is None else cy
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float32) + 0.5,
                          torch.arange(W, dtype=torch.float32) + 0.5, indexing="ij")

    def cam_dir(ii, jj):
        return torch.stack([(ii - cx) / focal, (jj - cy) / focal, torch.ones_like(ii)], -1)

    R = c2w[:3, :3]
    d0 = cam_dir(i, j) @ R.T
    dx = cam_dir(i + 1, j) @ R.T
    dy = cam_dir(i, j + 1) @ R.T
    radii = 0.5 * ((dx - d0).norm(dim=-1) + (dy - d0).norm(dim=-1)) * (2 / math.sqrt(12))
    d = d0 / d0.norm(dim=-1, keepdim=True)
    o = c2w[:3, 3].expand_as(d)
    parts = [o.reshape(-1, 3), d.reshape(-1, 3)]
    if cols == 7:
        parts.append(radii.reshape(-1, 1))
    return torch.cat(parts, -1).contiguous()


def sphere_occupancy(aabb, res=200, radius=1.0, holes_seed=None):
    """{0,1} occupancy on a res^3 lattice spanning `aabb`, 1 inside |x|<radius; volume[z][y][x]."""
    res3 = (res, res, res) if isinstance(res, int) else tuple(res)
    xs = [torch.linspace(float(aabb[0][a]), float(aabb[1][a]), res3[a]) for a in range(3)]
    zz, yy, xx = torch.meshgrid(xs[2], xs[1], xs[0], indexing="ij")
    vol = ((xx * xx + yy * yy + zz * zz) < radius * radius).float()
    if holes_seed is not None:
        g = torch.Generator().manual_seed(holes_seed)
        vol = vol * (torch.rand(vol.shape, generator=g) > 0.3).float()
    return orc.OccupancyGrid(aabb=aabb.clone(), volume=vol.contiguous())


def make_field(grid, aabb=None, density_shift=0.0, near_far=(2.0, 6.0), seed=SEED, occupancy="sphere",
               occ_res=200, holes_seed=None):
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3]) if aabb is None else aabb
    torch.manual_seed(seed)
    fld = orc.init_field(aabb, grid, density_shift=density_shift, near_far=list(near_far), step_ratio=0.5,
                         distance_scale=25.0, weight_thres=1e-4, **MODEL_KW)
    if occupancy == "sphere":
        r = float((aabb[1] - aabb[0]).min()) / 3  # 1.0 for the +-1.5 cube
        fld.occupancy = sphere_occupancy(aabb, occ_res, radius=r, holes_seed=holes_seed)
    return fld


# ---- the five BASELINE configs ------------------------------------------------
def config1(density_shift=0.0, occupancy="sphere", cols=6, holes_seed=1):
    """128^3, 100x100 rays from (0,0,4) looking -z (S=440)."""
    fld = make_field([128] * 3, density_shift=density_shift, occupancy=occupancy, holes_seed=holes_seed)
    c2w = look_at_c2w((0.0, 0.0, 4.0), up=(0.0, 1.0, 0.0))
    rays = pinhole_rays(100, 100, 100.0 / (2 * math.tan(0.5 * 0.6911112)), c2w, cols=cols)
    return fld, rays


def config2_rays(H=800, W=800, theta_deg=35.0, phi_deg=30.0):
    focal = 0.5 * W / math.tan(0.5 * 0.6911112)
    return pinhole_rays(H, W, focal, orbit_pose(theta_deg, phi_deg), cols=7)


def config2(density_shift=0.0, H=800, W=800):
    """300^3 lego-shaped, 800x800 Blender-style 7-col rays (S=1036)."""
    return make_field([300] * 3, density_shift=density_shift), config2_rays(H, W)


TRUCK_AABB = [[-1.35, -1.10, -0.55], [1.32, 1.14, 1.12]]


def config4(H=1080, W=1920):
    """Non-cubic T&T-like aabb, near_far=[0.01,6], 1920x1080."""
    aabb = torch.tensor(TRUCK_AABB)
    grid = orc.n_to_reso(300 ** 3, aabb)
    fld = make_field(grid, aabb=aabb, near_far=(0.01, 6.0), occ_res=(180, 200, 160))
    c2w = look_at_c2w((2.2, 1.6, 0.9), target=(0.0, 0.0, 0.25))
    return fld, pinhole_rays(H, W, 0.9 * W, c2w, cols=7)


def subsample(rays, n, seed=0):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(rays.shape[0], generator=g)[:n].sort().values
    return rays[idx].contiguous(), idx


def param_checksum(fld):
    """Cheap RNG-drift detector stored in the goldens."""
    return torch.stack([p.double().sum() for p in fld.params()]).numpy()
