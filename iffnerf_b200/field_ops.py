"""The callers on either side of the render path (SURVEY.md §8f "next" rows), as a mixin of TensorVMSplit.

Everything that touches factor or occupancy data runs in CUDA kernels behind the C ABI (include/tvm_b200.h):

  point / short-ray queries   compute_alpha, compute_densityfeature, compute_appfeature, the `sample_point_color`
                              sampler (pose_estimation/sampling.py:138,172,237-251)        csrc/query.cu, csrc/march.cu
  occupancy rebuild           getDenseAlpha / updateAlphaMask (models/tensorBase.py:643-696)  tvm_dense_alpha_mask
  ray filtering               filtering_rays (models/tensorBase.py:698-748)                   tvm_rays_hit_box /
                                                                                              tvm_sample_mask(ANYWHERE)
  factor resize / crop        upsample_volume_grid, shrink (models/tensoRF.py:258-316)        tvm_resize_factor

The host side only does the 3-element index arithmetic of `shrink` (on CPU tensors, so the voxel range is bit-identical
to the reference's) and replaces the Parameter objects like the reference does.  The regularisers on raw factors
(models/tensoRF.py:182-214) are a few torch reductions that keep adding into `.grad`.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib

MAT_MODE = [[0, 1], [0, 2], [1, 2]]
VEC_MODE = [2, 1, 0]


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FieldOpsMixin:
    # ------------------------------------------------------------------ samplers (API compatibility)
    def _box_mask(self, pts):
        box = self.aabb.to(pts.device)
        return ((pts >= box[0]) & (pts <= box[1])).all(dim=-1)

    def sample_ray(self, rays_o, rays_d, radii, is_train=True, N_samples=-1):
        """Signature and results of models/tensorBase.py:494-536 (points, z values, in-box mask) as tensor ops, for
        callers that want the samples themselves; the render kernels generate them internally."""
        n = N_samples if N_samples > 0 else self.nSamples
        box = self.aabb.to(rays_o.device)
        safe_d = torch.where(rays_d == 0, torch.full_like(rays_d, 1e-6), rays_d)
        entry = torch.minimum((box[1] - rays_o) / safe_d, (box[0] - rays_o) / safe_d).amax(-1)
        entry = entry.clamp(min=self.near_far[0], max=self.near_far[1])
        steps = torch.arange(n, dtype=rays_o.dtype, device=rays_o.device)
        if is_train:
            steps = steps.repeat(rays_d.shape[-2], 1)
            steps += torch.rand_like(steps[:, [0]])
        z = entry[..., None] + torch.multiply(self.stepSize.to(rays_o.device), steps)
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
        return pts, z, self._box_mask(pts)

    def sample_point_color(self, rays_o, rays_d, radii, N_samples=20, **kwargs):
        """models/tensorBase.py:623-638: N samples centred on each origin.  Passing this bound method as `sample_func`
        to forward() selects the kernels' TVM_F_POINT_SAMPLES sampler; calling it returns the reference's tensors."""
        half = N_samples // 2
        offsets = torch.arange(-half, N_samples - half, dtype=rays_o.dtype, device=rays_o.device)[None]
        z = self.stepSize.to(rays_o.device) * offsets
        pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., None]
        return pts, z, self._box_mask(pts)

    # ------------------------------------------------------------------ point queries
    def _flat_points(self, pts):
        if not pts.is_cuda:
            raise _lib.TvmError("point queries run on CUDA tensors only (no CPU path)")
        return pts.detach().reshape(-1, 3).float().contiguous()

    def _point_density(self, pts, mode, length=1.0):
        p = self._flat_points(pts)
        out = torch.empty((p.shape[0],), device=p.device)
        d, keep = self.field_desc()
        with torch.cuda.device(p.device):
            _lib.check(_lib.load().tvm_point_density(C.byref(d), _lib.ptr(p), p.shape[0], mode, float(length),
                                                     _lib.ptr(out), _stream(p.device)), "tvm_point_density")
        return out.view(pts.shape[:-1])

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
        p = self._flat_points(xyz_sampled)
        out = torch.empty((p.shape[0], self.app_dim), device=p.device)
        d, keep = self.field_desc()
        with torch.cuda.device(p.device):
            _lib.check(_lib.load().tvm_point_appfeature(C.byref(d), _lib.ptr(p), p.shape[0], _lib.ptr(out),
                                                        _stream(p.device)), "tvm_point_appfeature")
        return out.view(*xyz_sampled.shape[:-1], self.app_dim)

    def feature2density(self, density_features):
        if self.fea2denseAct == "softplus":
            return F.softplus(density_features + self.density_shift)
        return F.relu(density_features)

    # ------------------------------------------------------------------ occupancy rebuild
    def _param_device(self):
        return self.basis_mat.weight.device

    @staticmethod
    def _lattice_axes(grid, dev):
        # generated on the CPU like the reference's torch.linspace(0, 1, g) (tensorBase.py:649-653): same bits
        return [torch.linspace(0, 1, int(g)).to(dev) for g in grid]

    @torch.no_grad()
    def getDenseAlpha(self, gridSize=None):
        """models/tensorBase.py:643-665: (alpha [gx,gy,gz], dense_xyz [gx,gy,gz,3]) on a lattice spanning the aabb."""
        grid = self.gridSize.tolist() if gridSize is None else [int(g) for g in gridSize]
        dev = self._param_device()
        ax = self._lattice_axes(grid, dev)
        s = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1)
        box = self.aabb.to(dev)
        dense_xyz = box[0] * (1 - s) + box[1] * s
        alpha = self.compute_alpha(dense_xyz.view(-1, 3), self.stepSize.item()).view(dense_xyz.shape[:-1])
        return alpha, dense_xyz

    @torch.no_grad()
    def updateAlphaMask(self, gridSize=(200, 200, 200)):
        """models/tensorBase.py:667-696: rebuild the occupancy volume from the density field (dense alpha, 3x3x3
        max-pool, threshold) and return the tight box of the occupied lattice points — one C-ABI call
        (tvm_dense_alpha_mask: lattice, query, pooling, threshold and box reduction on the device)."""
        from .tensorf import AlphaGridMask
        gx, gy, gz = (int(g) for g in gridSize)
        dev = self._param_device()
        ax = self._lattice_axes((gx, gy, gz), dev)
        lib = _lib.load()
        volume = torch.empty((gz, gy, gx), dtype=torch.float32, device=dev)
        box = torch.empty((7,), dtype=torch.float32, device=dev)
        ws = torch.empty((lib.tvm_dense_alpha_workspace_bytes(gx, gy, gz),), dtype=torch.uint8, device=dev)
        d, keep = self.field_desc()
        with torch.cuda.device(dev):
            _lib.check(lib.tvm_dense_alpha_mask(C.byref(d), _lib.ptr(ax[0]), _lib.ptr(ax[1]), _lib.ptr(ax[2]), gx, gy, gz,
                                                float(self.stepSize.item()), float(self.alphaMask_thres),
                                                _lib.ptr(volume), _lib.ptr(box), _lib.ptr(ws), ws.numel(), _stream(dev)),
                       "tvm_dense_alpha_mask")
        host = box.cpu()
        if host[6] == 0:
            raise RuntimeError("updateAlphaMask: no lattice point passes alphaMask_thres (empty field)")
        self.last_alpha_rest = float(host[6]) / float(gx * gy * gz)     # the 'alpha rest %' the reference prints
        self.alphaMask = AlphaGridMask(dev, self.aabb.to(dev), volume, contraction_type=self.contraction_type)
        return box[:6].view(2, 3).clone()

    # ------------------------------------------------------------------ ray filtering
    @torch.no_grad()
    def filtering_rays(self, all_rays, all_rgbs, N_samples=256, chunk=10240 * 5, bbox_only=False):
        """models/tensorBase.py:698-748: keep the rays whose slab interval against the aabb is non-empty (bbox_only) or
        that have a sample inside an occupied cell of the alphaMask (tested on every sample, also outside the aabb,
        like the reference).  `chunk` only bounds the upload size; the kernels take any number of rays."""
        dev = self._param_device()
        flat = all_rays.reshape(-1, all_rays.shape[-1])
        if not bbox_only and self.alphaMask is None:
            raise RuntimeError("filtering_rays(bbox_only=False) needs an alphaMask")
        lib = _lib.load()
        keep_parts = []
        piece = max(int(chunk), 1 << 20)
        for a in range(0, flat.shape[0], piece):
            rays = flat[a:a + piece].to(dev).float().contiguous()
            if bbox_only:
                hit = torch.empty((rays.shape[0],), dtype=torch.uint8, device=dev)
                d, keep = self.field_desc(need_params=False)
                with torch.cuda.device(dev):
                    _lib.check(lib.tvm_rays_hit_box(C.byref(d), _lib.ptr(rays), rays.shape[0], rays.shape[1],
                                                    _lib.ptr(hit), _stream(dev)), "tvm_rays_hit_box")
                keep_parts.append(hit.bool().cpu())
            else:
                _, counts = self.sample_mask(rays, N_samples=N_samples, want_bits=False, anywhere=True)
                keep_parts.append((counts > 0).cpu())
        mask = torch.cat(keep_parts).view(all_rgbs.shape[:-1])
        return all_rays[mask], all_rgbs[mask]

    # ------------------------------------------------------------------ factor resize / crop
    def _resized(self, factor, h2, w2, mode=0, y_off=0, x_off=0):
        """New Parameter holding `factor` ([1,C,H,W]) resampled (mode 0, bilinear align_corners=True) or cropped."""
        src = factor.data.contiguous()
        _, c, h, w = src.shape
        dst = torch.empty((1, c, h2, w2), dtype=src.dtype, device=src.device)
        if not src.is_cuda:
            raise _lib.TvmError("factor resizing runs on CUDA tensors only (no CPU path)")
        with torch.cuda.device(src.device):
            _lib.check(_lib.load().tvm_resize_factor(_lib.ptr(src), c, h, w, _lib.ptr(dst), h2, w2, mode, y_off, x_off,
                                                     _stream(src.device)), "tvm_resize_factor")
        return torch.nn.Parameter(dst)

    @torch.no_grad()
    def up_sampling_VM(self, plane_coef, line_coef, res_target):
        """models/tensoRF.py:258-270: every plane to (res[m1], res[m0]), every line to (res[v], 1)."""
        for k, ((m0, m1), v) in enumerate(zip(MAT_MODE, VEC_MODE)):
            plane_coef[k] = self._resized(plane_coef[k], int(res_target[m1]), int(res_target[m0]))
            line_coef[k] = self._resized(line_coef[k], int(res_target[v]), 1)
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
        """models/tensoRF.py:280-316: crop the factors to the voxel range that covers new_aabb; when the alphaMask
        lattice differs from the field grid the box is snapped to the kept voxels.  The voxel range is computed on CPU
        tensors with the reference's fp32 arithmetic (it decides array shapes)."""
        box = self.aabb.detach().cpu().float()
        want = new_aabb.detach().cpu().float()
        units = self.units.detach().cpu().float()
        grid = self.gridSize.detach().cpu()
        lo = torch.round(torch.round((want[0] - box[0]) / units)).long()
        hi = torch.minimum(torch.round((want[1] - box[0]) / units).long() + 1, grid)
        lo_l, hi_l = lo.tolist(), hi.tolist()
        for k, ((m0, m1), v) in enumerate(zip(MAT_MODE, VEC_MODE)):
            for lines, planes in ((self.density_line, self.density_plane), (self.app_line, self.app_plane)):
                lines[k] = self._resized(lines[k], hi_l[v] - lo_l[v], 1, mode=1, y_off=lo_l[v])
                planes[k] = self._resized(planes[k], hi_l[m1] - lo_l[m1], hi_l[m0] - lo_l[m0], mode=1,
                                          y_off=lo_l[m1], x_off=lo_l[m0])
        mask_grid = None if self.alphaMask is None else self.alphaMask.gridSize.detach().cpu()
        if mask_grid is None or not torch.all(mask_grid == grid):
            frac_lo, frac_hi = lo / (grid - 1), (hi - 1) / (grid - 1)
            want = torch.stack(((1 - frac_lo) * box[0] + frac_lo * box[1], (1 - frac_hi) * box[0] + frac_hi * box[1]))
        self.aabb = want.to(self.aabb.device)
        self.update_stepSize((hi - lo).tolist())

    # ------------------------------------------------------------------ regularisers (models/tensoRF.py:182-214)
    @staticmethod
    def vectorDiffs(vector_comps):
        """Mean |<v_i, v_j>|, i != j, of the line components of each mode (orthogonality regulariser)."""
        total = 0
        for comp in vector_comps:
            v = comp.view(comp.shape[1], comp.shape[2])
            gram = v @ v.t()
            n = gram.shape[0]
            total = total + (gram.abs().sum() - gram.diagonal().abs().sum()) / (n * (n - 1))
        return total

    def vector_comp_diffs(self):
        return self.vectorDiffs(self.density_line) + self.vectorDiffs(self.app_line)

    def density_L1(self):
        return sum(p.abs().mean() for p in self.density_plane) + sum(p.abs().mean() for p in self.density_line)

    def TV_loss_density(self, reg):
        return sum(reg(p) for p in self.density_plane) * 1e-2

    def TV_loss_app(self, reg):
        return sum(reg(p) for p in self.app_plane) * 1e-2
