"""The callers on either side of the render path (SURVEY.md §8f "next" rows), as a mixin of TensorVMSplit:

  * point / short-ray queries: `sample_point_color` sampler, `compute_alpha`, `compute_densityfeature`
    (pose_estimation/sampling.py:138,172,237-251) — CUDA kernels behind the C ABI;
  * grid maintenance: `getDenseAlpha`, `updateAlphaMask`, `filtering_rays`, `upsample_volume_grid`, `shrink`
    (models/tensorBase.py:643-748, models/tensoRF.py:258-316; run <= 7 times per 30 k iterations) — the dense
    density evaluation and the ray/occupancy tests run on the kernels, the rest is tensor bookkeeping;
  * the regularisers on raw factors (models/tensoRF.py:182-214) — tiny torch ops that keep adding into `.grad`.
"""
from __future__ import annotations

import ctypes as C
import time

import torch
import torch.nn.functional as F

from . import _lib

MAT_MODE = [[0, 1], [0, 2], [1, 2]]
VEC_MODE = [2, 1, 0]


class FieldOpsMixin:
    # ------------------------------------------------------------------ samplers (API compatibility)
    def sample_ray(self, rays_o, rays_d, radii, is_train=True, N_samples=-1):
        """models/tensorBase.py:494-536 as tensor ops (the render kernels do this internally; callers such as
        filtering_rays only need the API)."""
        S = N_samples if N_samples > 0 else self.nSamples
        near, far = self.near_far
        aabb = self.aabb.to(rays_o.device)
        vec = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        t_min = torch.minimum((aabb[1] - rays_o) / vec, (aabb[0] - rays_o) / vec).amax(-1).clamp(min=near, max=far)
        rng = torch.arange(S, dtype=rays_o.dtype, device=rays_o.device)
        if is_train:
            rng = rng.repeat(rays_d.shape[-2], 1)
            rng += torch.rand_like(rng[:, [0]])
        z = t_min[..., None] + torch.multiply(self.stepSize.to(rays_o.device), rng)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
        outside = ((aabb[0] > pts) | (pts > aabb[1])).any(dim=-1)
        return pts, z, ~outside

    def sample_point_color(self, rays_o, rays_d, radii, N_samples=20, **kwargs):
        """models/tensorBase.py:623-638.  Passing this bound method as `sample_func` to forward() selects the
        kernels' TVM_F_POINT_SAMPLES sampler; calling it directly returns the same tensors as the reference."""
        before = N_samples // 2
        rng = torch.arange(-before, N_samples - before, dtype=rays_o.dtype, device=rays_o.device)[None]
        step = self.stepSize.to(rays_o.device) * rng
        pts = rays_o[..., None, :] + rays_d[..., None, :] * step[..., None]
        aabb = self.aabb.to(rays_o.device)
        outside = ((aabb[0] > pts) | (pts > aabb[1])).any(dim=-1)
        return pts, step, ~outside

    # ------------------------------------------------------------------ point queries
    def _point_density(self, pts, mode, length=1.0):
        from .tensorf import _stream
        if not pts.is_cuda:
            raise _lib.TvmError("point queries run on CUDA tensors only (no CPU path)")
        shape = pts.shape[:-1]
        p = pts.detach().reshape(-1, 3).float().contiguous()
        out = torch.empty((p.shape[0],), device=p.device)
        d, keep = self.field_desc()
        _lib.check(_lib.load().tvm_point_density(C.byref(d), _lib.ptr(p), p.shape[0], mode, float(length),
                                                 _lib.ptr(out), _stream(p.device)), "tvm_point_density")
        return out.view(shape)

    @torch.no_grad()
    def compute_densityfeature(self, xyz_sampled):
        """models/tensoRF.py:216-235: raw sigma feature at NORMALISED coordinates [M,3] (inference only)."""
        return self._point_density(xyz_sampled, 0)

    @torch.no_grad()
    def compute_alpha(self, xyz_locs, length=1):
        """models/tensorBase.py:756-773: 1 - exp(-sigma*length) at world points, gated by the alphaMask."""
        return self._point_density(xyz_locs, 1, float(length))

    @torch.no_grad()
    def compute_appfeature(self, xyz_sampled):
        """models/tensoRF.py:237-256: basis_mat(app_plane (x) app_line) at NORMALISED coordinates [M,3] -> [M, app_dim]
        (inference only; pose_estimation/sampling.py:535-541 feeds it to Ref.compute_normals)."""
        from .tensorf import _stream
        if not xyz_sampled.is_cuda:
            raise _lib.TvmError("point queries run on CUDA tensors only (no CPU path)")
        shape = xyz_sampled.shape[:-1]
        p = xyz_sampled.detach().reshape(-1, 3).float().contiguous()
        out = torch.empty((p.shape[0], self.app_dim), device=p.device)
        d, keep = self.field_desc()
        _lib.check(_lib.load().tvm_point_appfeature(C.byref(d), _lib.ptr(p), p.shape[0], _lib.ptr(out),
                                                    _stream(p.device)), "tvm_point_appfeature")
        return out.view(*shape, self.app_dim)

    def feature2density(self, density_features):
        if self.fea2denseAct == "softplus":
            return F.softplus(density_features + self.density_shift)
        return F.relu(density_features)

    # ------------------------------------------------------------------ grid maintenance
    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None):
        """models/tensorBase.py:643-665: alpha on a dense lattice spanning the aabb (one kernel launch)."""
        gridSize = self.gridSize.tolist() if gridSize is None else [int(g) for g in gridSize]
        dev = self.basis_mat.weight.device
        samples = torch.stack(torch.meshgrid(torch.linspace(0, 1, gridSize[0]), torch.linspace(0, 1, gridSize[1]),
                                             torch.linspace(0, 1, gridSize[2]), indexing="ij"), -1).to(dev)
        aabb = self.aabb.to(dev)
        dense_xyz = aabb[0] * (1 - samples) + aabb[1] * samples
        alpha = self.compute_alpha(dense_xyz.view(-1, 3), self.stepSize.item()).view(dense_xyz.shape[:-1])
        return alpha, dense_xyz

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200)):
        """models/tensorBase.py:667-696: rebuild the occupancy volume (3x3x3 max-pool, threshold) and return the
        tight aabb of the occupied region."""
        from .tensorf import AlphaGridMask
        gridSize = [int(g) for g in gridSize]
        alpha, dense_xyz = self.getDenseAlpha(gridSize)
        dense_xyz = dense_xyz.transpose(0, 2).contiguous()
        alpha = alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None]
        alpha = F.max_pool3d(alpha, kernel_size=3, padding=1, stride=1).view(gridSize[::-1])
        alpha = (alpha >= self.alphaMask_thres).float()
        dev = alpha.device
        self.alphaMask = AlphaGridMask(dev, self.aabb.to(dev), alpha, contraction_type=self.contraction_type)
        valid_xyz = dense_xyz[alpha > 0.5]
        return torch.stack((valid_xyz.amin(0), valid_xyz.amax(0)))

    @torch.no_grad()
    def filtering_rays(self, all_rays, all_rgbs, N_samples=256, chunk=10240 * 5, bbox_only=False):
        """models/tensorBase.py:698-748: keep the rays that hit the box (bbox_only) or the occupancy volume."""
        dev = self.basis_mat.weight.device
        flat = all_rays.reshape(-1, all_rays.shape[-1])
        masks = []
        big = max(int(chunk), 1 << 20)           # the kernels do not need small chunks
        for a in range(0, flat.shape[0], big):
            rays = flat[a:a + big].to(dev)
            if bbox_only:
                o, d = rays[..., :3], rays[..., 3:6]
                aabb = self.aabb.to(dev)
                vec = torch.where(d == 0, torch.full_like(d, 1e-6), d)
                ra, rb = (aabb[1] - o) / vec, (aabb[0] - o) / vec
                keep = torch.maximum(ra, rb).amin(-1) > torch.minimum(ra, rb).amax(-1)
            else:
                if self.alphaMask is None:
                    raise RuntimeError("filtering_rays(bbox_only=False) needs an alphaMask")
                # in-aabb AND occupied somewhere along the ray == the kernels' ray_valid count > 0; the reference
                # tests the occupancy of every sample incl. those outside the aabb, which the aabb ⊂ mask-aabb
                # invariant makes equivalent for any ray that can contribute
                _, counts = self.sample_mask(rays, N_samples=N_samples, want_bits=False)
                keep = counts > 0
            masks.append(keep.cpu())
        mask = torch.cat(masks).view(all_rgbs.shape[:-1])
        return all_rays[mask], all_rgbs[mask]

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        """models/tensoRF.py:258-270."""
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            plane_coef[k] = torch.nn.Parameter(F.interpolate(plane_coef[k].data, size=(res_target[m1], res_target[m0]),
                                                             mode="bilinear", align_corners=True))
            line_coef[k] = torch.nn.Parameter(F.interpolate(line_coef[k].data, size=(res_target[VEC_MODE[k]], 1),
                                                            mode="bilinear", align_corners=True))
        return plane_coef, line_coef

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        """models/tensoRF.py:272-278."""
        res_target = [int(r) for r in res_target]
        self.app_plane, self.app_line = self.up_sampling_VM(self.app_plane, self.app_line, res_target)
        self.density_plane, self.density_line = self.up_sampling_VM(self.density_plane, self.density_line, res_target)
        self.update_stepSize(res_target)

    @torch.no_grad()
    def shrink(self, new_aabb):
        """models/tensoRF.py:280-316: crop the factors to the voxel range covering new_aabb."""
        dev = self.basis_mat.weight.device
        aabb = self.aabb.to(dev)
        new_aabb = new_aabb.to(dev)
        units = self.units.to(dev)
        grid = self.gridSize.to(dev)
        t_l, b_r = (new_aabb[0] - aabb[0]) / units, (new_aabb[1] - aabb[0]) / units
        t_l, b_r = torch.round(torch.round(t_l)).long(), torch.round(b_r).long() + 1
        b_r = torch.stack([b_r, grid]).amin(0)
        for k in range(3):
            v = VEC_MODE[k]
            self.density_line[k] = torch.nn.Parameter(self.density_line[k].data[..., t_l[v]:b_r[v], :].contiguous())
            self.app_line[k] = torch.nn.Parameter(self.app_line[k].data[..., t_l[v]:b_r[v], :].contiguous())
            m0, m1 = MAT_MODE[k]
            self.density_plane[k] = torch.nn.Parameter(
                self.density_plane[k].data[..., t_l[m1]:b_r[m1], t_l[m0]:b_r[m0]].contiguous())
            self.app_plane[k] = torch.nn.Parameter(
                self.app_plane[k].data[..., t_l[m1]:b_r[m1], t_l[m0]:b_r[m0]].contiguous())
        if self.alphaMask is None or not torch.all(self.alphaMask.gridSize.to(dev) == grid):
            t_l_r, b_r_r = t_l / (grid - 1), (b_r - 1) / (grid - 1)
            correct = torch.zeros_like(new_aabb)
            correct[0] = (1 - t_l_r) * aabb[0] + t_l_r * aabb[1]
            correct[1] = (1 - b_r_r) * aabb[0] + b_r_r * aabb[1]
            new_aabb = correct
        new_size = (b_r - t_l).tolist()
        self.aabb = new_aabb
        self.update_stepSize(new_size)

    # ------------------------------------------------------------------ regularisers (models/tensoRF.py:182-214)
    @staticmethod
    def vectorDiffs(vector_comps):
        total = 0
        for comp in vector_comps:
            n_comp, n_size = comp.shape[1:-1]
            v = comp.view(n_comp, n_size)
            dotp = torch.matmul(v, v.transpose(-1, -2))
            off_diag = dotp.view(-1)[1:].view(n_comp - 1, n_comp + 1)[..., :-1]
            total = total + torch.mean(torch.abs(off_diag))
        return total

    def vector_comp_diffs(self):
        return self.vectorDiffs(self.density_line) + self.vectorDiffs(self.app_line)

    def density_L1(self):
        total = 0
        for k in range(3):
            total = total + torch.mean(torch.abs(self.density_plane[k])) + torch.mean(torch.abs(self.density_line[k]))
        return total

    def TV_loss_density(self, reg):
        total = 0
        for k in range(3):
            total = total + reg(self.density_plane[k]) * 1e-2
        return total

    def TV_loss_app(self, reg):
        total = 0
        for k in range(3):
            total = total + reg(self.app_plane[k]) * 1e-2
        return total
