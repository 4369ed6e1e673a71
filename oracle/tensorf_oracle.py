"""TEST INFRASTRUCTURE ONLY — CPU oracle for the TensoRF-VM ray-render path.

This is a restatement (not a copy) of the reference's algorithm for the path
SURVEY.md §8 scopes, written as plain functions over a parameter record.  It
issues the same ATen fp32 op sequence as the reference so that, on CPU, it is
*bit-identical* to the unmodified reference (`oracle/make_golden.py` asserts
this in the authoring container, which is what pins the oracle; the committed
`tests/golden/*.npz` are outputs of the reference itself).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this file.  The product package
(`iffnerf_b200/`) never does; it fails loudly when its CUDA library is missing.

Citations are `file:line` in the reference tree (mbortolon97/IFFNeRF).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.nn.functional as F

MAT_MODE = ((0, 1), (0, 2), (1, 2))  # models/tensorBase.py:311
VEC_MODE = (2, 1, 0)                 # models/tensorBase.py:312


# --------------------------------------------------------------------------- #
# parameter record
# --------------------------------------------------------------------------- #
@dataclass
class OccupancyGrid:
    """AlphaGridMask (models/tensorBase.py:50-64): {0,1} fp32 volume [Dz,Dy,Dx] with its own aabb."""
    aabb: torch.Tensor          # [2,3]
    volume: torch.Tensor        # [Dz,Dy,Dx] float32

    @property
    def inv_size(self):         # models/tensorBase.py:59  (1.0 / aabbSize * 2)
        return 1.0 / (self.aabb[1] - self.aabb[0]) * 2


@dataclass
class Field:
    aabb: torch.Tensor                      # [2,3] fp32
    grid: List[int]                         # gridSize (x,y,z)
    density_plane: List[torch.Tensor]       # 3 x [1,Cs,G[m1],G[m0]]
    density_line: List[torch.Tensor]        # 3 x [1,Cs,G[v],1]
    app_plane: List[torch.Tensor]           # 3 x [1,Ca,G[m1],G[m0]]
    app_line: List[torch.Tensor]            # 3 x [1,Ca,G[v],1]
    basis: torch.Tensor                     # [app_dim, sum(Ca)]
    mlp_w: List[torch.Tensor]               # [featureC,in], [featureC,featureC], [3,featureC]
    mlp_b: List[torch.Tensor]
    near_far: List[float] = field(default_factory=lambda: [2.0, 6.0])
    step_ratio: float = 0.5
    density_shift: float = -10.0
    distance_scale: float = 25.0
    weight_thres: float = 1e-4
    fea2dense: str = "softplus"
    view_pe: int = 2
    fea_pe: int = 2
    occupancy: Optional[OccupancyGrid] = None

    def params(self):
        return (list(self.density_plane) + list(self.density_line) + list(self.app_plane)
                + list(self.app_line) + [self.basis] + list(self.mlp_w) + list(self.mlp_b))


def step_geometry(aabb: torch.Tensor, grid, step_ratio: float):
    """models/tensorBase.py:354-375 (aabb contraction branch). All fp32 CPU tensor ops."""
    aabb = aabb.detach().cpu()
    size = aabb[1] - aabb[0]
    inv = 2.0 / size
    g = torch.tensor([int(v) for v in grid], dtype=torch.long)
    units = size / (g - 1)
    step = torch.mean(units) * step_ratio
    diag = torch.sqrt(torch.sum(torch.square(size)))
    n_samples = int((diag / step).item()) + 1
    return {"aabbSize": size, "invaabbSize": inv, "units": units, "stepSize": step,
            "aabbDiag": diag, "nSamples": n_samples}


def n_to_reso(n_voxels: int, bbox: torch.Tensor):
    """utils.py:20-24."""
    lo, hi = bbox
    vox = ((hi - lo).prod() / n_voxels).pow(1 / len(lo))
    return ((hi - lo) / vox).long().tolist()


def init_field(aabb, grid, *, n_sigma=(16, 16, 16), n_app=(48, 48, 48), app_dim=27, feature_c=128,
               view_pe=2, fea_pe=2, scale=0.1, **scalars) -> Field:
    """Random-init field drawing from torch's global RNG in the SAME order as the reference
    constructor (models/tensoRF.py:155-170, then basis Linear, then MLPRender_Fea
    models/tensorBase.py:165-183), so one seed yields identical parameters on both sides."""
    def one_svd(n_comp):
        planes, lines = [], []
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            planes.append(scale * torch.randn((1, n_comp[k], grid[m1], grid[m0])))
            lines.append(scale * torch.randn((1, n_comp[k], grid[VEC_MODE[k]], 1)))
        return planes, lines
    dp, dl = one_svd(n_sigma)
    ap, al = one_svd(n_app)
    basis = torch.nn.Linear(sum(n_app), app_dim, bias=False)
    in_c = 2 * view_pe * 3 + 2 * fea_pe * app_dim + 3 + app_dim
    l1 = torch.nn.Linear(in_c, feature_c)
    l2 = torch.nn.Linear(feature_c, feature_c)
    l3 = torch.nn.Linear(feature_c, 3)
    torch.nn.init.constant_(l3.bias, 0)
    return Field(aabb=aabb, grid=[int(v) for v in grid], density_plane=dp, density_line=dl, app_plane=ap,
                 app_line=al, basis=basis.weight.detach().clone(),
                 mlp_w=[l.weight.detach().clone() for l in (l1, l2, l3)],
                 mlp_b=[l.bias.detach().clone() for l in (l1, l2, l3)],
                 view_pe=view_pe, fea_pe=fea_pe, **scalars)


# --------------------------------------------------------------------------- #
# stages
# --------------------------------------------------------------------------- #
def sample_along_rays(fld: Field, o, d, n_samples: int, jitter=None):
    """models/tensorBase.py:494-536.  jitter: None (eval) or [N,1] U[0,1) (train, :507-509)."""
    geo = step_geometry(fld.aabb, fld.grid, fld.step_ratio)
    S = n_samples if n_samples > 0 else geo["nSamples"]
    near, far = fld.near_far
    safe = torch.where(d == 0, torch.full_like(d, 1e-6), d)
    ra = (fld.aabb[1] - o) / safe
    rb = (fld.aabb[0] - o) / safe
    t0 = torch.minimum(ra, rb).amax(-1).clamp(min=near, max=far)
    idx = torch.arange(S, dtype=o.dtype)
    if jitter is not None:
        idx = idx.repeat(d.shape[-2], 1)
        idx += jitter
    offs = torch.multiply(geo["stepSize"], idx)
    z = t0[..., None] + offs
    pts = o[..., None, :] + d[..., None, :] * z[..., None]
    outside = ((fld.aabb[0] > pts) | (pts > fld.aabb[1])).any(dim=-1)
    return pts, z, ~outside


def sample_around_points(fld: Field, o, d, n_samples: int):
    """sample_point_color, models/tensorBase.py:623-638: n samples centred on the origin, z = step*(i - n//2)."""
    geo = step_geometry(fld.aabb, fld.grid, fld.step_ratio)
    before = n_samples // 2
    idx = torch.arange(-before, n_samples - before, dtype=o.dtype)[None]
    z = geo["stepSize"] * idx
    pts = o[..., None, :] + d[..., None, :] * z[..., None]
    outside = ((fld.aabb[0] > pts) | (pts > fld.aabb[1])).any(dim=-1)
    return pts, z, ~outside


def point_alpha(fld: Field, pts, length=1.0):
    """TensorBase.compute_alpha, models/tensorBase.py:756-773."""
    if fld.occupancy is not None:
        keep = occupancy_value(fld.occupancy, pts) > 0
    else:
        keep = torch.ones_like(pts[:, 0], dtype=bool)
    sigma = torch.zeros(pts.shape[:-1])
    if keep.any():
        sigma[keep] = to_density(fld, density_feature(fld, normalize(fld, pts[keep])))
    return 1 - torch.exp(-sigma * length).view(pts.shape[:-1])


def dense_alpha_volume(fld: Field, grid=(200, 200, 200), thres=1e-4):
    """getDenseAlpha + updateAlphaMask, models/tensorBase.py:643-696 -> ({0,1} volume [Dz,Dy,Dx], tight aabb)."""
    geo = step_geometry(fld.aabb, fld.grid, fld.step_ratio)
    samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, grid[0]), torch.linspace(0, 1, grid[1]),
                                         torch.linspace(0, 1, grid[2]), indexing="ij"), -1)
    dense = fld.aabb[0] * (1 - samples) + fld.aabb[1] * samples
    alpha = torch.zeros_like(dense[..., 0])
    for i in range(grid[0]):
        alpha[i] = point_alpha(fld, dense[i].view(-1, 3), geo["stepSize"]).view((grid[1], grid[2]))
    dense = dense.transpose(0, 2).contiguous()
    alpha = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
    alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1).view(list(grid)[::-1])
    vol = (alpha >= thres).float()
    valid = dense[vol > 0.5]
    return vol, torch.stack((valid.amin(0), valid.amax(0)))


def occupancy_value(occ: OccupancyGrid, pts):
    """models/tensorBase.py:66-83: trilinear grid_sample of the {0,1} volume, align_corners=True."""
    n = (pts - occ.aabb[0]) * occ.inv_size - 1
    vol = occ.volume.view(1, 1, *occ.volume.shape[-3:])
    return F.grid_sample(vol, n.view(1, -1, 1, 1, 3), align_corners=True).view(-1)


def normalize(fld: Field, pts):
    """models/tensorBase.py:389-397 (aabb branch)."""
    geo = step_geometry(fld.aabb, fld.grid, fld.step_ratio)
    return (pts - fld.aabb[0]) * geo["invaabbSize"] - 1


def _vm_coords(p):
    plane = torch.stack([p[..., list(MAT_MODE[k])] for k in range(3)]).view(3, -1, 1, 2)
    line = torch.stack([p[..., VEC_MODE[k]] for k in range(3)])
    line = torch.stack((torch.zeros_like(line), line), dim=-1).view(3, -1, 1, 2)
    return plane, line


def density_feature(fld: Field, p):
    """models/tensoRF.py:216-235."""
    cp, cl = _vm_coords(p)
    out = torch.zeros((p.shape[0],))
    for k in range(3):
        pv = F.grid_sample(fld.density_plane[k], cp[[k]], align_corners=True).view(-1, p.shape[0])
        lv = F.grid_sample(fld.density_line[k], cl[[k]], align_corners=True).view(-1, p.shape[0])
        out = out + torch.sum(pv * lv, dim=0)
    return out


def app_feature(fld: Field, p):
    """models/tensoRF.py:237-256 (incl. basis_mat, :158)."""
    cp, cl = _vm_coords(p)
    pv, lv = [], []
    for k in range(3):
        pv.append(F.grid_sample(fld.app_plane[k], cp[[k]], align_corners=True).view(-1, p.shape[0]))
        lv.append(F.grid_sample(fld.app_line[k], cl[[k]], align_corners=True).view(-1, p.shape[0]))
    prod = (torch.cat(pv) * torch.cat(lv)).T
    return F.linear(prod, fld.basis)


def to_density(fld: Field, f):
    """models/tensorBase.py:750-754."""
    if fld.fea2dense == "softplus":
        return F.softplus(f + fld.density_shift)
    return F.relu(f)


def composite_weights(sigma, dist):
    """models/tensorBase.py:23-35."""
    alpha = 1.0 - torch.exp(-sigma * dist)
    trans = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), 1.0 - alpha + 1e-10], -1), -1)
    return alpha, alpha * trans[:, :-1]


def freq_encode(x, n_freq):
    """models/tensorBase.py:14-20."""
    bands = 2 ** torch.arange(n_freq).float()
    y = (x[..., None] * bands).reshape(x.shape[:-1] + (n_freq * x.shape[-1],))
    return torch.cat([torch.sin(y), torch.cos(y)], dim=-1)


def shade(fld: Field, viewdirs, feat):
    """MLPRender_Fea.forward, models/tensorBase.py:185-195."""
    cols = [feat, viewdirs]
    if fld.fea_pe > 0:
        cols.append(freq_encode(feat, fld.fea_pe))
    if fld.view_pe > 0:
        cols.append(freq_encode(viewdirs, fld.view_pe))
    h = torch.cat(cols, dim=-1)
    h = torch.relu(F.linear(h, fld.mlp_w[0], fld.mlp_b[0]))
    h = torch.relu(F.linear(h, fld.mlp_w[1], fld.mlp_b[1]))
    return torch.sigmoid(F.linear(h, fld.mlp_w[2], fld.mlp_b[2]))


# --------------------------------------------------------------------------- #
# the path
# --------------------------------------------------------------------------- #
def render_chunk(fld: Field, rays, white_bg=False, bg_color=None, n_samples=-1, jitter=None, point_samples=False):
    """TensorBase.forward, models/tensorBase.py:775-917 (aabb contraction, sample_ray branch).

    Returns a dict with the reference's 6 outputs plus the two masks."""
    view = rays[:, 3:6]
    if point_samples:                                               # sample_func=sample_point_color (:792-803)
        pts, z, valid = sample_around_points(fld, rays[:, :3], view, n_samples)
    else:
        pts, z, valid = sample_along_rays(fld, rays[:, :3], view, n_samples, jitter)
    dists = torch.cat((z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])), dim=-1)
    if fld.occupancy is not None:                                   # :832-837
        keep = occupancy_value(fld.occupancy, pts[valid]) > 0
        bad = ~valid
        bad[valid] |= ~keep
        valid = ~bad
    sigma = torch.zeros(pts.shape[:-1])
    if valid.any():                                                 # :841-846
        pts = normalize(fld, pts)
        sigma[valid] = to_density(fld, density_feature(fld, pts[valid]))
    alpha, w = composite_weights(sigma, dists * fld.distance_scale)  # :849
    app = w > fld.weight_thres                                      # :851
    feats = torch.zeros((*pts.shape[:2], fld.basis.shape[0]))       # :872-878
    if app.any():
        feats[app] = app_feature(fld, pts[app])
    lit = app.any(dim=-1)                                           # :886-896
    acc = torch.sum(w, -1)
    ray_feat = torch.sum(w[..., None] * feats, -2)
    rgb = torch.zeros((*view.shape[:-1], 3))
    rgb[lit] = shade(fld, view[lit], ray_feat[lit])
    if bg_color is None:                                            # :898-904
        bg_color = torch.ones(3) if white_bg else torch.zeros(3)
    rgb = (rgb * acc[..., None] + bg_color * (1.0 - acc[..., None])).clamp(0, 1)
    with torch.no_grad():                                           # :906-908
        depth = torch.sum(w * z, -1) + (1.0 - acc) * rays[..., -1]
    return {"rgb_map": rgb, "depth_map": depth, "acc_map": acc, "alpha": alpha, "z_vals": z,
            "dists": dists, "ray_valid": valid, "app_mask": app, "weight": w, "ray_feat": ray_feat}


def render_rays(fld: Field, rays, chunk=4096, n_samples=-1, white_bg=None, bg_color=None, jitter=None,
                keys=("rgb_map", "depth_map")):
    """OctreeRender_trilinear_fast, renderer.py:12-25 (chunk loop + concatenation); extra keys on request."""
    parts = {k: [] for k in keys}
    for a in range(0, rays.shape[0], chunk):
        out = render_chunk(fld, rays[a:a + chunk], white_bg=white_bg, bg_color=bg_color, n_samples=n_samples,
                           jitter=None if jitter is None else jitter[a:a + chunk])
        for k in keys:
            parts[k].append(out[k])
    return {k: torch.cat(v) for k, v in parts.items()}


def pack_valid_bits(valid: torch.Tensor):
    """[N,S] bool -> [N, ceil(S/32)] int32 words, bit (i%32) of word (i//32) = sample i (little-endian)."""
    import numpy as np
    v = valid.cpu().numpy().astype(np.uint8)
    n, s = v.shape
    words = (s + 31) // 32
    pad = np.zeros((n, words * 32), dtype=np.uint8)
    pad[:, :s] = v
    packed = np.packbits(pad.reshape(n, words, 32), axis=-1, bitorder="little")
    return torch.from_numpy(packed.reshape(n, words, 4).copy().view("<u4").reshape(n, words).astype(np.int64))


def train_loss(out, target):
    """train.py:293 + train.py:328-329 (MSE + 0.1*mean(exp(|alpha|)))."""
    return torch.mean((out["rgb_map"] - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(out["alpha"])))


# --------------------------------------------------------------------------- #
# ray generation
# --------------------------------------------------------------------------- #
def camera_directions(H: int, W: int, K: torch.Tensor):
    """Pixel-centre directions K^-1 [x+.5, y+.5, 1] for every pixel and for its +1-x / +1-y neighbours
    (get_ray_directions_Ks, ray_utils.py:28-60): returns three [1,H,W,3] grids."""
    xs = torch.arange(W, dtype=torch.float32) + 0.5
    ys = torch.arange(H, dtype=torch.float32) + 0.5
    gx, gy = torch.meshgrid(xs, ys, indexing="xy")                     # [H,W] each
    Kinv = torch.inverse(K.reshape(-1, 3, 3))                          # ray_utils.py:50
    grids = []
    for ox, oy in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0)):
        pix = torch.stack([gx + ox, gy + oy, torch.ones_like(gx)], 0).reshape(1, 3, -1)
        grids.append((Kinv @ pix).reshape(-1, 3, H, W).permute(0, 2, 3, 1))
    return grids


def pixel_rays(K: torch.Tensor, c2w: torch.Tensor, pixels: torch.Tensor, H: int, W: int, renormalize: bool = True):
    """[N,7] rays of `pixels` [N,2] (x,y) through pose `c2w` [3|4,4]: the chain of
    inerf/estimate_pose_inerf.py:96-99 (unit view directions), ray_utils.py:63-100 (rotate, origin, radii) and
    :149-164 (pixel indexing, F.normalize, cat).  Differentiable w.r.t. c2w."""
    ori, dx, dy = camera_directions(H, W, K)
    view = ori / torch.linalg.norm(ori, dim=-1, keepdim=True)
    R = c2w[..., :3, :3]

    def rot(v):                                                       # (v[..., None, :] * R).sum(-1), ray_utils.py:76
        return (v[..., None, :] * R).sum(-1)
    d = rot(view)                                                     # same creation order as ray_utils.py:76-83
    wx, wy = rot(dx), rot(dy)
    wo = rot(ori)
    o = c2w[..., :3, 3].unsqueeze(-2).expand(d.shape)
    radii = (0.5 * (torch.linalg.norm(wx - wo, dim=-1) + torch.linalg.norm(wy - wo, dim=-1))[..., None]) \
        * (2 / math.sqrt(12))                                         # ray_utils.py:90-98
    px, py = pixels[:, 0].long(), pixels[:, 1].long()
    d = d[0, py, px]
    if renormalize:
        d = F.normalize(d, p=2, dim=-1)                               # estimate_pose_inerf.py:159
    return torch.cat((o[0, py, px], d, radii[0, py, px]), dim=-1)
