"""Host-side mirror of the reference's radiance-field module for the TensoRF-VM render path.

Same class names, constructor kwargs, `state_dict` keys and return contracts as the reference
(`TensorVMSplit`, models/tensoRF.py:151-316, on top of `TensorBase`, models/tensorBase.py:262-917),
but `forward` is one call into the hand-written sm_100a kernels behind the C ABI of
include/tvm_b200.h — there is no CPU or eager-PyTorch fallback for the render path.

Parameters stay in the reference layout ([1,C,H,W] planes, [1,C,L,1] lines) so checkpoints, Adam,
`upsample_volume_grid` and `shrink` keep working; a channel-last packed shadow, keyed on
(data_ptr, _version) of every factor, is what the kernels read.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .field_ops import FieldOpsMixin

MAT_MODE = [[0, 1], [0, 2], [1, 2]]
VEC_MODE = [2, 1, 0]


def positional_encoding(positions, freqs):
    """models/tensorBase.py:14-20 (torch; used by the autograd shade path and by callers)."""
    bands = (2 ** torch.arange(freqs, device=positions.device)).float()
    scaled = (positions[..., None] * bands).reshape(positions.shape[:-1] + (freqs * positions.shape[-1],))
    return torch.cat([torch.sin(scaled), torch.cos(scaled)], dim=-1)


class AlphaGridMask(torch.nn.Module):
    """Occupancy volume with its own aabb (models/tensorBase.py:50-83).

    `alpha_volume` keeps the reference's [1,1,Dz,Dy,Dx] fp32 layout; the render kernels read a
    derived one-byte-per-cell corner code (`cells()`), rebuilt when the volume changes."""

    def __init__(self, device, aabb, alpha_volume, contraction_type="aabb"):
        super().__init__()
        if contraction_type != "aabb":
            raise NotImplementedError("only the aabb contraction is on the B200 render path")
        self.device = device
        self.contraction_type = contraction_type
        self.aabb = aabb.to(device)
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.invgridSize = 1.0 / self.aabbSize * 2
        self.alpha_volume = alpha_volume.view(1, 1, *alpha_volume.shape[-3:])
        self.gridSize = torch.tensor([alpha_volume.shape[-1], alpha_volume.shape[-2], alpha_volume.shape[-3]],
                                     dtype=torch.int64).to(device)
        self._cells = None
        self._cells_key = None
        # host copies of the fp32 scalars the kernel needs (computed with the reference's op order on CPU)
        a = self.aabb.detach().cpu().float()
        self._lo = a[0].tolist()
        self._inv = (1.0 / (a[1] - a[0]) * 2).tolist()

    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb[0]) * self.invgridSize - 1

    def sample_alpha(self, xyz_sampled):
        """Point query (models/tensorBase.py:66-72); not on the ray-render hot path."""
        n = self.normalize_coord(xyz_sampled)
        return F.grid_sample(self.alpha_volume, n.view(1, -1, 1, 1, 3), align_corners=True).view(-1)

    def cells(self):
        vol = self.alpha_volume
        key = (vol.data_ptr(), vol._version, tuple(vol.shape))
        if self._cells is None or self._cells_key != key:
            if not vol.is_cuda:
                raise _lib.TvmError("AlphaGridMask volume must live on a CUDA device for rendering")
            v = vol.detach().reshape(vol.shape[-3:]).contiguous().float()
            dz, dy, dx = v.shape
            lib = _lib.load()
            cells = torch.empty((lib.tvm_occupancy_bytes(dx, dy, dz),), dtype=torch.uint8, device=v.device)
            with torch.cuda.device(v.device):
                _lib.check(lib.tvm_pack_occupancy(_lib.ptr(v), dx, dy, dz, _lib.ptr(cells), _stream(v.device)),
                           "tvm_pack_occupancy")
            self._cells, self._cells_key = cells, key
            self._cells_dims = (dx, dy, dz)
            self._coarse_off = int(lib.tvm_occupancy_coarse_offset(dx, dy, dz))
        return self._cells


class MLPRender_Fea(torch.nn.Module):
    """Shading head (models/tensorBase.py:165-195); same parameter names (`mlp.0/2/4`)."""

    def __init__(self, inChanel, viewpe=6, feape=6, featureC=128):
        super().__init__()
        self.in_mlpC = 2 * viewpe * 3 + 2 * feape * inChanel + 3 + inChanel
        self.viewpe, self.feape = viewpe, feape
        first = torch.nn.Linear(self.in_mlpC, featureC)
        second = torch.nn.Linear(featureC, featureC)
        last = torch.nn.Linear(featureC, 3)
        self.mlp = torch.nn.Sequential(first, torch.nn.ReLU(inplace=True), second, torch.nn.ReLU(inplace=True), last)
        torch.nn.init.constant_(self.mlp[-1].bias, 0)

    def forward(self, pts, viewdirs, features, *args):
        cols = [features, viewdirs]
        if self.feape > 0:
            cols.append(positional_encoding(features, self.feape))
        if self.viewpe > 0:
            cols.append(positional_encoding(viewdirs, self.viewpe))
        return torch.sigmoid(self.mlp(torch.cat(cols, dim=-1))), None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class TensorVMSplit(FieldOpsMixin, torch.nn.Module):
    """VM-decomposed radiance field with a B200-native renderer.

    Constructor signature and defaults follow TensorBase.__init__ (models/tensorBase.py:262-326)."""

    # launch-size cap for the chunk-free eval path (rays per kernel launch)
    max_launch_rays = 1 << 20
    # eval-mode forward(): also return the per-sample tensors alpha / z_vals / dists [N,S] like the reference (True).
    # False returns None for them; rays then terminate at T < early_term_eps in forward AND backward, which is what the
    # pose-refinement loop wants (inerf/estimate_pose_inerf.py:166 reads rgb and opacity only).
    eval_sample_outputs = True
    grad_forward_tc3 = True      # differentiable forward with a frozen MLP_Fea head (pose mode): tensor-core shading kernel
    # `Ref` head: fused tail kernels in both directions.  False = run the head as torch ops — an explicit cross-check
    # switch for tests / scripts; there is no silent fallback: a head without a fused kernel raises while these are True
    ref_kernel_train = True
    ref_kernel = True
    # transmittance below which an eval ray stops marching (error on rgb/acc <= this value)
    early_term_eps = 1e-5
    # shading-stage arithmetic on the no-grad path:
    #   "fp32"  SIMT FFMA kernel (rgb within 1e-4 of the reference)
    #   "tc3"   tcgen05 tensor-core kernel, bf16x3 split operands + fp32 accumulate (fp32-equivalent: ~1e-6 of "fp32")
    #   "bf16"  tcgen05 tensor-core kernel, plain bf16 operands (rgb within 1e-2 — BASELINE.json's "bf16 MLP mode")
    #   "auto"  (default) "tc3" when the head's shape fits that kernel (fea_pe = view_pe = 2 does), else "fp32"
    mlp_precision = "auto"
    # eval renders without per-sample outputs: two-kernel march (a 64-register sigma-march that emits per-ray appearance
    # sample lists + an appearance-gather kernel with a register texel cache) instead of the fused one-kernel march
    split_app = True

    def __init__(self, aabb, gridSize, device, density_n_comp=8, appearance_n_comp=24, app_dim=27,
                 shadingMode="MLP_PE", alphaMask=None, near_far=[2.0, 6.0], density_shift=-10,
                 alphaMask_thres=0.001, distance_scale=25, rayMarch_weight_thres=0.0001, pos_pe=6, view_pe=6,
                 fea_pe=6, featureC=128, step_ratio=2.0, fea2denseAct="softplus", contraction_type="aabb",
                 step_size_bg=0.1):
        super().__init__()
        if contraction_type != "aabb":
            raise NotImplementedError("contraction_type='unisphere' is outside the B200 render path (SURVEY.md 8a)")
        if shadingMode not in ("MLP_Fea", "Ref"):
            raise NotImplementedError(f"shadingMode={shadingMode!r}: only 'MLP_Fea' and 'Ref' are on the B200 render path")
        if fea2denseAct not in ("softplus", "relu"):
            raise ValueError(f"unknown fea2denseAct {fea2denseAct!r}")
        if isinstance(density_n_comp, int):
            density_n_comp = [density_n_comp] * 3
        if isinstance(appearance_n_comp, int):
            appearance_n_comp = [appearance_n_comp] * 3
        self.density_n_comp = list(density_n_comp)
        self.app_n_comp = list(appearance_n_comp)
        self.app_dim = app_dim
        self.aabb = aabb
        self.alphaMask = alphaMask
        self.device = device
        self.density_shift = density_shift
        self.alphaMask_thres = alphaMask_thres
        self.distance_scale = distance_scale
        self.rayMarch_weight_thres = rayMarch_weight_thres
        self.fea2denseAct = fea2denseAct
        self.near_far = near_far
        self.step_ratio = step_ratio
        self.contraction_type = contraction_type
        self.step_size_bg = step_size_bg
        self.matMode = MAT_MODE
        self.vecMode = VEC_MODE
        self.comp_w = [1, 1, 1]
        self.update_stepSize(gridSize)
        self.init_svd_volume(gridSize[0], device)
        self.shadingMode, self.pos_pe, self.view_pe, self.fea_pe, self.featureC = \
            shadingMode, pos_pe, view_pe, fea_pe, featureC
        if shadingMode == "MLP_Fea":
            self.renderModule = MLPRender_Fea(self.app_dim, view_pe, fea_pe, featureC).to(device)
        else:       # "Ref": march stage on the kernels, the ~4 kFLOP/ray head as torch ops (ref_head.py)
            from .ref_head import Ref
            self.renderModule = Ref(self.app_dim, viewpe=view_pe, feature_c=featureC).to(device)
        self.native_shade = shadingMode == "MLP_Fea"
        self.it = 0
        self._packed = None
        self._packed_key = None
        self._mlp_packed = None
        self._mlp_key = None
        self._mlp_tc = None
        self._mlp_tc_key = None
        self._bg_cache = {}
        self.grad_sync = None      # sharding.GradSync when training data-parallel

    # ------------------------------------------------------------------ geometry
    def update_stepSize(self, gridSize):
        """models/tensorBase.py:354-375.  The fp32 scalars are evaluated on CPU tensors with the
        reference's op order so they are bit-identical to the CPU reference whatever the device."""
        grid = [int(g) for g in gridSize]
        box = self.aabb.detach().cpu().float()
        size = box[1] - box[0]
        inv = 2.0 / size
        g = torch.tensor(grid, dtype=torch.long)
        units = size / (g - 1)
        step = torch.mean(units) * self.step_ratio
        diag = torch.sqrt(torch.sum(torch.square(size)))
        dev = self.device
        self.aabbSize = size.to(dev)
        self.invaabbSize = inv.to(dev)
        self.gridSize = g.to(dev)
        self.units = units.to(dev)
        self.stepSize = step.to(dev)
        self.aabbDiag = diag.to(dev)
        self.nSamples = int((diag / step).item()) + 1
        self.n_samples_bg = 0
        self._host = {"aabb": box.reshape(-1).tolist(), "inv": inv.tolist(), "grid": grid, "step": float(step.item())}
        self._packed_key = None

    # ------------------------------------------------------------------ parameters
    def init_svd_volume(self, res, device):
        """models/tensoRF.py:155-158; RNG draw order identical to the reference (one seed -> same factors)."""
        self.density_plane, self.density_line = self.init_one_svd(self.density_n_comp, self._host["grid"], 0.1, device)
        self.app_plane, self.app_line = self.init_one_svd(self.app_n_comp, self._host["grid"], 0.1, device)
        self.basis_mat = torch.nn.Linear(sum(self.app_n_comp), self.app_dim, bias=False).to(device)

    def init_one_svd(self, n_component, gridSize, scale, device):
        planes, lines = [], []
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            planes.append(torch.nn.Parameter(scale * torch.randn((1, n_component[k], gridSize[m1], gridSize[m0]))))
            lines.append(torch.nn.Parameter(scale * torch.randn((1, n_component[k], gridSize[VEC_MODE[k]], 1))))
        return torch.nn.ParameterList(planes).to(device), torch.nn.ParameterList(lines).to(device)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        """models/tensoRF.py:172-180."""
        groups = [{"params": self.density_line, "lr": lr_init_spatialxyz},
                  {"params": self.density_plane, "lr": lr_init_spatialxyz},
                  {"params": self.app_line, "lr": lr_init_spatialxyz},
                  {"params": self.app_plane, "lr": lr_init_spatialxyz},
                  {"params": self.basis_mat.parameters(), "lr": lr_init_network}]
        if isinstance(self.renderModule, torch.nn.Module):
            groups.append({"params": self.renderModule.parameters(), "lr": lr_init_network})
        return groups

    def get_kwargs(self):
        """models/tensorBase.py:402-422 (same keys, so `eval(model_name)(**kwargs)` round-trips)."""
        return {"aabb": self.aabb, "gridSize": self.gridSize.tolist(), "density_n_comp": self.density_n_comp,
                "appearance_n_comp": self.app_n_comp, "app_dim": self.app_dim,
                "contraction_type": self.contraction_type, "density_shift": self.density_shift,
                "alphaMask_thres": self.alphaMask_thres, "distance_scale": self.distance_scale,
                "rayMarch_weight_thres": self.rayMarch_weight_thres, "fea2denseAct": self.fea2denseAct,
                "near_far": self.near_far, "step_ratio": self.step_ratio, "shadingMode": self.shadingMode,
                "pos_pe": self.pos_pe, "view_pe": self.view_pe, "fea_pe": self.fea_pe, "featureC": self.featureC}

    def save(self, path):
        """models/tensorBase.py:424-442: same `.th` dictionary (packbits occupancy)."""
        ckpt = {"model_name": type(self).__name__, "kwargs": self.get_kwargs(), "state_dict": self.state_dict()}
        if self.alphaMask is not None:
            vol = self.alphaMask.alpha_volume.bool().cpu().numpy()
            ckpt["alphaMask.shape"] = vol.shape
            ckpt["alphaMask.mask"] = np.packbits(vol.reshape(-1))
            ckpt["alphaMask.aabb"] = self.alphaMask.aabb.cpu()
        torch.save(ckpt, path)

    def load(self, ckpt):
        """models/tensorBase.py:444-458."""
        if "alphaMask.aabb" in ckpt.keys():
            n = int(np.prod(ckpt["alphaMask.shape"]))
            vol = torch.from_numpy(np.unpackbits(ckpt["alphaMask.mask"])[:n].reshape(ckpt["alphaMask.shape"]))
            self.alphaMask = AlphaGridMask(self.device, ckpt["alphaMask.aabb"].to(self.device),
                                           vol.float().to(self.device), contraction_type=self.contraction_type)
        self.load_state_dict(ckpt["state_dict"])

    def normalize_coord(self, xyz_sampled):
        return (xyz_sampled - self.aabb[0].to(xyz_sampled.device)) * self.invaabbSize - 1

    # ------------------------------------------------------------------ kernel-side views of the parameters
    def _factor_params(self):
        return (list(self.density_plane) + list(self.app_plane), list(self.density_line) + list(self.app_line))

    def _factor_layout(self):
        """Float offsets of the 12 packed sections (each 64-float aligned).  Plane rows are padded to an odd number of
        texels (tvm_plane_pitch in csrc/tvm_math.cuh: L1 bank spreading of vertically adjacent texels)."""
        g = self._host["grid"]
        off, out = 0, {"dplane": [], "dline": [], "aplane": [], "aline": []}

        def take(n):
            nonlocal off
            start = off
            off = (off + n + 63) // 64 * 64
            return start
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            out["dplane"].append(take((g[m0] | 1) * g[m1] * self.density_n_comp[k]))   # rows padded to an odd pitch
            out["dline"].append(take(g[VEC_MODE[k]] * self.density_n_comp[k]))
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            out["aplane"].append(take((g[m0] | 1) * g[m1] * self.app_n_comp[k]))
            out["aline"].append(take(g[VEC_MODE[k]] * self.app_n_comp[k]))
        out["total"] = off
        return out

    def _check_shapes(self):
        g = self._host["grid"]
        for k in range(3):
            m0, m1 = MAT_MODE[k]
            exp_p = (g[m1], g[m0])
            exp_l = (g[VEC_MODE[k]], 1)
            for name, plist, llist in (("density", self.density_plane, self.density_line),
                                       ("app", self.app_plane, self.app_line)):
                if tuple(plist[k].shape[-2:]) != exp_p or tuple(llist[k].shape[-2:]) != exp_l:
                    raise _lib.TvmError(f"{name} factor {k} shape {tuple(plist[k].shape)}/{tuple(llist[k].shape)} "
                                        f"does not match gridSize {g}; call update_stepSize after resizing")

    def _base_desc(self):
        h = self._host
        d = _lib.FieldDesc()
        d.aabb[:] = h["aabb"]
        d.inv_aabb[:] = h["inv"]
        d.grid[:] = h["grid"]
        d.step_size = h["step"]
        d.near_t, d.far_t = float(self.near_far[0]), float(self.near_far[1])
        d.density_shift = float(self.density_shift)
        d.distance_scale = float(self.distance_scale)
        d.weight_thres = float(self.rayMarch_weight_thres)
        d.early_term_eps = float(self.early_term_eps)
        d.act = 0 if self.fea2denseAct == "softplus" else 1
        d.n_sigma[:] = self.density_n_comp
        d.n_app[:] = self.app_n_comp
        d.app_dim = self.app_dim
        d.fea_pe, d.view_pe, d.feature_c = self.fea_pe, self.view_pe, self.featureC
        lay = self._factor_layout()
        d.dplane_off[:] = lay["dplane"]
        d.dline_off[:] = lay["dline"]
        d.aplane_off[:] = lay["aplane"]
        d.aline_off[:] = lay["aline"]
        d.n_factor_floats = lay["total"]
        return d

    def grad_workspace(self, device):
        """Persistent flat fp32 gradient workspace of the backward kernels: [packed factor gradients | packed MLP
        gradients | basis_mat gradient].  Zero on entry of every backward (the unpack kernels clear what they read), so
        no per-step 70 MB memset / allocation, and data-parallel training all-reduces it as ONE bucket."""
        d = self._base_desc()
        lib = _lib.load()
        nf, nm = int(d.n_factor_floats), int(lib.tvm_mlp_grad_floats(C.byref(d))) if self.native_shade else 0
        nb = self.basis_mat.weight.numel()
        sync = self.grad_sync if hasattr(self.grad_sync, "alloc_workspace") else None
        key = (nf, nm, nb, str(device), id(sync))
        if getattr(self, "_grad_ws_key", None) != key:
            # data parallel: the sync object decides where the workspace lives (NVLink peer memory when it can; the
            # allocation is then a collective, which is fine: every rank reaches its first backward together)
            self._grad_ws = (sync.alloc_workspace(nf + nm + nb, device) if sync is not None
                             else torch.zeros(nf + nm + nb, dtype=torch.float32, device=device))
            self._grad_ws_key = key
        w = self._grad_ws
        return w, w[:nf], w[nf:nf + nm], w[nf + nm:].view_as(self.basis_mat.weight)

    def invalidate_packed(self):
        """Forget the cache keys of every packed parameter shadow: the next use re-packs from the parameters.  (Needed
        after in-place updates that bypass autograd version counters, e.g. an optimiser step replayed by a CUDA graph.)"""
        self._packed_key = self._mlp_key = self._mlp_tc_key = None
        self._ref_key = None

    def packed_factors(self):
        """Channel-last shadow of the 12 factor tensors, re-packed only when a factor changed."""
        planes, lines = self._factor_params()
        key = tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in planes + lines)
        if self._packed is None or self._packed_key != key:
            self._check_shapes()
            dev = planes[0].device
            if dev.type != "cuda":
                raise _lib.TvmError("TensorVMSplit parameters must live on a CUDA device for rendering "
                                    "(there is no CPU path)")
            d = self._base_desc()
            if self._packed is None or self._packed.numel() != d.n_factor_floats or self._packed.device != dev:
                self._packed = torch.zeros(int(d.n_factor_floats), dtype=torch.float32, device=dev)
            lib = _lib.load()
            srcs_p = [p.detach().contiguous() for p in planes]
            srcs_l = [p.detach().contiguous() for p in lines]
            with torch.cuda.device(dev):
                _lib.check(lib.tvm_pack_factors(C.byref(d), _lib.ptr_array(srcs_p), _lib.ptr_array(srcs_l),
                                                _lib.ptr(self._packed), _stream(dev)), "tvm_pack_factors")
            self._packed_key = key
        return self._packed

    def packed_mlp(self):
        mods = [self.renderModule.mlp[i] for i in (0, 2, 4)]
        ps = [m.weight for m in mods] + [m.bias for m in mods]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if self._mlp_packed is None or self._mlp_key != key:
            dev = ps[0].device
            d = self._base_desc()
            lib = _lib.load()
            n = lib.tvm_mlp_pack_floats(C.byref(d))
            if self._mlp_packed is None or self._mlp_packed.numel() != n or self._mlp_packed.device != dev:
                self._mlp_packed = torch.zeros(int(n), dtype=torch.float32, device=dev)
            w1, w2, w3, b1, b2, b3 = [p.detach().contiguous() for p in ps]
            with torch.cuda.device(dev):
                _lib.check(lib.tvm_pack_mlp(C.byref(d), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                            _lib.ptr(w3), _lib.ptr(b3), _lib.ptr(self._mlp_packed), _stream(dev)),
                           "tvm_pack_mlp")
            self._mlp_key = key
        return self._mlp_packed

    def _shade_mode(self):
        """Resolves mlp_precision to the shading kernel actually used: 'fp32' | 'tc3' | 'bf16'."""
        mode = self.mlp_precision
        if mode not in ("auto", "fp32", "tc3", "bf16"):
            raise ValueError(f"mlp_precision must be 'auto', 'fp32', 'tc3' or 'bf16', got {mode!r}")
        if mode != "auto":
            return mode
        key = (self.app_dim, self.fea_pe, self.view_pe, self.featureC, tuple(self.app_n_comp))
        if getattr(self, "_auto_key", None) != key:
            d = self._base_desc()
            ok = _lib.load().tvm_mlp_tc3_supported(C.byref(d)) != 0
            self._auto_key, self._auto_mode = key, ("tc3" if ok else "fp32")
        return self._auto_mode

    def packed_mlp_tc(self, split=False):
        """bf16 operand images of basis_mat + MLP weights for the tcgen05 shade kernels (split: hi|lo pairs)."""
        mods = [self.renderModule.mlp[i] for i in (0, 2, 4)]
        ps = [self.basis_mat.weight] + [m.weight for m in mods]
        key = (split,) + tuple((p.data_ptr(), p._version) for p in ps)
        if self._mlp_tc is None or self._mlp_tc_key != key:
            dev = ps[0].device
            d = self._base_desc()
            lib = _lib.load()
            size_fn = lib.tvm_mlp_tc3_pack_bytes if split else lib.tvm_mlp_tc_pack_bytes
            pack_fn = lib.tvm_pack_mlp_tc3 if split else lib.tvm_pack_mlp_tc
            n = size_fn(C.byref(d))
            if self._mlp_tc is None or self._mlp_tc.numel() != n or self._mlp_tc.device != dev:
                self._mlp_tc = torch.zeros(int(n), dtype=torch.uint8, device=dev)
            b, w1, w2, w3 = [p.detach().contiguous() for p in ps]
            with torch.cuda.device(dev):
                _lib.check(pack_fn(C.byref(d), _lib.ptr(b), _lib.ptr(w1), _lib.ptr(w2), _lib.ptr(w3),
                                   _lib.ptr(self._mlp_tc), _stream(dev)), "tvm_pack_mlp_tc")
            self._mlp_tc_key = key
        return self._mlp_tc

    def packed_ref_head(self):
        """`tvm_ref_head` descriptor + packed parameter buffer of the `Ref` head for tvm_shade_ref_fwd (layout in
        include/tvm_b200.h); None when the head's configuration has no fused kernel (the torch tail is used)."""
        rm = self.renderModule
        if not getattr(rm, "predicted_normals", False) or self.app_dim != 27:
            return None
        lins = {"normal": rm.normal_mlp[0], "tint": rm.tint_color_mlp[0], "rough": rm.roughness_mlp[0],
                "diffuse": rm.diffuse_color_mlp[0], "bott": rm.bottleneck_mlp, "spec": rm.specular_mlp[0]}
        ps = [p for l in lins.values() for p in (l.weight, l.bias)] + [rm.dir_enc_fn.mat]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if getattr(self, "_ref_key", None) != key:
            dev = ps[0].device
            h = _lib.RefHead()
            ml = rm.dir_enc_fn.ml_array.detach().cpu().long()
            h.in_c, h.feature_c = self.app_dim, lins["bott"].out_features
            h.n_pairs, h.l_max = int(ml.shape[1]), int(rm.dir_enc_fn.mat.shape[0]) - 1
            if h.n_pairs > 32 or lins["spec"].in_features != h.feature_c + 2 * h.n_pairs + 1:
                return None
            h.m[:h.n_pairs] = ml[0].tolist()
            h.l[:h.n_pairs] = ml[1].tolist()
            pre = rm.rgb_premultiplier if abs(rm.rgb_premultiplier - 1.0) > 1e-7 else 1.0
            bias = rm.rgb_bias if rm.rgb_bias > 1e-7 else 0.0
            h.rgb_premultiplier, h.rgb_bias, h.rgb_padding = pre, bias, float(rm.rgb_padding)
            h.diffuse_shift, h.rough_shift = float(rm.diffuse_shift), float(rm.rough_shift)
            lib = _lib.load()
            offs = (C.c_int32 * 8)()
            _lib.check(lib.tvm_ref_head_layout(C.byref(h), offs), "tvm_ref_head_layout")
            small_w, bott_w, small_b, bott_b, spec_w, spec_b, ide, in4 = list(offs)
            buf = torch.zeros(int(lib.tvm_ref_head_floats(C.byref(h))), dtype=torch.float32, device=dev)
            small = torch.cat([lins[k].weight.detach() for k in ("normal", "tint", "rough", "diffuse")], 0)    # [10, in_c]
            buf[small_w:small_w + 10 * in4].view(10, in4)[:, :h.in_c] = small
            buf[bott_w:bott_w + h.feature_c * in4].view(h.feature_c, in4)[:, :h.in_c] = lins["bott"].weight.detach()
            buf[small_b:small_b + 10] = torch.cat([lins[k].bias.detach() for k in ("normal", "tint", "rough", "diffuse")])
            buf[bott_b:bott_b + h.feature_c] = lins["bott"].bias.detach()
            sw = lins["spec"].weight.detach().reshape(-1)
            buf[spec_w:spec_w + sw.numel()] = sw
            buf[spec_b:spec_b + 3] = lins["spec"].bias.detach()
            mat = rm.dir_enc_fn.mat.detach().float().reshape(-1)
            buf[ide:ide + mat.numel()] = mat
            h.params = buf.data_ptr()
            self._ref_key, self._ref_head = key, (h, buf)
        return self._ref_head

    def field_desc(self, need_params=True):
        """Full descriptor with device pointers; keeps the referenced tensors alive via the returned tuple."""
        d = self._base_desc()
        keep = []
        if need_params:
            pf = self.packed_factors()
            basis = self.basis_mat.weight.detach().contiguous()
            d.factors, d.basis = pf.data_ptr(), basis.data_ptr()
            keep += [pf, basis]
            if self.native_shade:
                pm = self.packed_mlp()
                d.mlp = pm.data_ptr()
                keep.append(pm)
            mode = self._shade_mode() if self.native_shade else "fp32"
            if self.native_shade and mode == "bf16":
                tc = self.packed_mlp_tc()
                d.mlp_tc = tc.data_ptr()
                keep.append(tc)
            elif self.native_shade and mode == "tc3":
                tc = self.packed_mlp_tc(split=True)
                d.mlp_tc3 = tc.data_ptr()
                keep.append(tc)

        if self.alphaMask is not None:
            cells = self.alphaMask.cells()
            dx, dy, dz = self.alphaMask._cells_dims
            d.occ_cells = cells.data_ptr()
            d.occ_dims[:] = [dx, dy, dz]
            d.occ_coarse = cells.data_ptr() + self.alphaMask._coarse_off
            d.occ_cdims[:] = [(dx + 15) // 16, (dy + 15) // 16, (dz + 15) // 16]
            d.occ_lo[:] = self.alphaMask._lo
            d.occ_inv[:] = self.alphaMask._inv
            keep.append(cells)
        return d, keep

    def _bg(self, bg_color, white_bg, device):
        if bg_color is not None:
            if bg_color.numel() != 3:
                raise ValueError(f"bg_color must hold 3 values (one colour for the batch), got shape {tuple(bg_color.shape)}")
            return bg_color.detach().to(device=device, dtype=torch.float32).reshape(3).contiguous()
        key = (bool(white_bg), str(device))
        if key not in self._bg_cache:
            self._bg_cache[key] = (torch.ones if white_bg else torch.zeros)(3, device=device)
        return self._bg_cache[key]

    @staticmethod
    def _prep_rays(rays):
        if not rays.is_cuda:
            raise _lib.TvmError("rays must be on a CUDA device (the render path has no CPU implementation)")
        if rays.dim() != 2 or rays.shape[1] < 6:
            raise ValueError(f"rays must be [N, 6|7], got {tuple(rays.shape)}")
        return rays.detach().float().contiguous()

    # ------------------------------------------------------------------ kernels
    @torch.no_grad()
    def sample_mask(self, rays, N_samples=-1, jitter=None, want_bits=True, point_samples=False, anywhere=False):
        """`ray_valid` of TensorBase.forward (sample_ray + aabb + alphaMask, tensorBase.py:820-837) as packed bits
        [N, ceil(S/32)] (int32 view of uint32 words) and per-ray counts.  Bit-exact w.r.t. the reference.
        `anywhere`: the occupancy test alone decides, also outside the aabb (filtering_rays, tensorBase.py:728-737)."""
        rays = self._prep_rays(rays)
        S = N_samples if N_samples > 0 else self.nSamples
        n = rays.shape[0]
        d, keep = self.field_desc(need_params=False)
        bits = torch.empty((n, (S + 31) // 32), dtype=torch.int32, device=rays.device) if want_bits else None
        counts = torch.empty((n,), dtype=torch.int32, device=rays.device)
        jit = None if jitter is None else jitter.detach().float().reshape(-1).contiguous()
        lib = _lib.load()
        with torch.cuda.device(rays.device):
            _lib.check(lib.tvm_sample_mask(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, _lib.ptr(jit),
                                           (_lib.F_POINT_SAMPLES if point_samples else 0) |
                                           (_lib.F_MASK_ANYWHERE if anywhere else 0), _lib.ptr(bits),
                                           _lib.ptr(counts), _stream(rays.device)), "tvm_sample_mask")
        return bits, counts

    @torch.no_grad()
    def render_eval(self, rays, N_samples=-1, white_bg=False, bg_color=None, jitter=None, sample_outputs=False,
                    early_term=True, want_counts=False, keep_workspace=False, out_rgb=None, out_depth=None,
                    point_samples=False, scatter=None):
        """One launch pair (march + shade) over `rays` [N,6|7] on the GPU; no autograd.

        Returns a dict with rgb_map [N,3], depth_map [N], acc_map [N] and, when `sample_outputs`, the
        reference's per-sample outputs alpha / z_vals / dists [N,S] (which disables early termination)."""
        rays = self._prep_rays(rays)
        with torch.cuda.device(rays.device):       # the C-ABI launches go to the thread's current device
            return self._render_eval(rays, N_samples, white_bg, bg_color, jitter, sample_outputs, early_term, want_counts,
                                     keep_workspace, out_rgb, out_depth, point_samples, scatter)

    def _render_eval(self, rays, N_samples, white_bg, bg_color, jitter, sample_outputs, early_term, want_counts,
                     keep_workspace, out_rgb, out_depth, point_samples, scatter=None):
        """scatter = (dst_index int64 [N] | None, [rgb base pointers], [depth base pointers]): the shading epilogue writes
        ray i at offset dst_index[i] of every destination image (tvm_scatter_out; sharding.render_sharded's peer
        placement) instead of into local rgb_map / depth_map tensors, which are then absent from the result."""
        dev = rays.device
        S = N_samples if N_samples > 0 else self.nSamples
        n = rays.shape[0]
        d, keep = self.field_desc()
        lib = _lib.load()
        if n == 0:
            out_rgb = out_depth = None
        if scatter is not None:
            out = {"rgb_map": None, "depth_map": None, "acc_map": torch.empty((n,), device=dev)}
        else:
            out = {"rgb_map": out_rgb if out_rgb is not None else torch.empty((n, 3), device=dev),
                   "depth_map": out_depth if out_depth is not None else torch.empty((n,), device=dev),
                   "acc_map": torch.empty((n,), device=dev)}
            assert out["rgb_map"].is_contiguous() and out["depth_map"].is_contiguous()
        alpha = z = dists = None
        if sample_outputs:
            alpha, z, dists = (torch.empty((n, S), device=dev) for _ in range(3))
            out.update(alpha=alpha, z_vals=z, dists=dists)
        vcount = acount = None
        if want_counts:
            vcount = torch.empty((n,), dtype=torch.int32, device=dev)
            acount = torch.empty((n,), dtype=torch.int32, device=dev)
            out.update(valid_count=vcount, app_count=acount)
        flags = _lib.F_EARLY_TERM if (early_term and not sample_outputs) else 0
        if self.split_app and not sample_outputs:
            flags |= _lib.F_SPLIT_APP
            head_kernel = self.native_shade or (self.ref_kernel and self.packed_ref_head() is not None)
            if keep_workspace or not head_kernel:
                flags |= _lib.F_ZERO_UNLIT          # someone reads the ray_feat rows of unlit rays
        need = C.c_size_t(0)
        _lib.check(lib.tvm_workspace_bytes(C.byref(d), n, flags & _lib.F_SPLIT_APP, C.byref(need)), "tvm_workspace_bytes")
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        shade_flags = 0
        if self.native_shade and self._shade_mode() == "bf16":
            shade_flags = _lib.F_MLP_BF16
        elif self.native_shade and self._shade_mode() == "tc3":
            shade_flags = _lib.F_MLP_TC3
        if not self.native_shade or scatter is not None:
            flags |= _lib.F_NO_SHADE
        elif self._shade_mode() == "bf16":
            flags |= _lib.F_MLP_BF16
        elif self._shade_mode() == "tc3":
            flags |= _lib.F_MLP_TC3
        if point_samples:
            flags |= _lib.F_POINT_SAMPLES
        jit = None if jitter is None else jitter.detach().to(dev).float().reshape(-1).contiguous()
        bg = self._bg(bg_color, white_bg, dev)
        _lib.check(lib.tvm_render_fwd(C.byref(d), _lib.ptr(rays), n, rays.shape[1], S, _lib.ptr(jit), _lib.ptr(bg),
                                      flags, _lib.ptr(out["rgb_map"]), _lib.ptr(out["depth_map"]),
                                      _lib.ptr(out["acc_map"]), _lib.ptr(alpha), _lib.ptr(z), _lib.ptr(dists),
                                      None, _lib.ptr(vcount), _lib.ptr(acount), _lib.ptr(ws), ws.numel(),
                                      _stream(dev)), "tvm_render_fwd")
        if scatter is not None and n > 0:
            idx, rgb_ptrs, depth_ptrs = scatter
            sc = _lib.ScatterOut()
            if idx is not None:
                idx = idx.to(device=dev, dtype=torch.int64).contiguous()
                sc.dst_index = idx.data_ptr()
            sc.n_dst = len(rgb_ptrs)
            for k, (pr, pd) in enumerate(zip(rgb_ptrs, depth_ptrs)):
                sc.rgb[k], sc.depth[k] = int(pr), int(pd)
            if self.native_shade:
                _lib.check(lib.tvm_shade_fwd_scatter(C.byref(d), _lib.ptr(rays), n, rays.shape[1], _lib.ptr(bg), shade_flags,
                                                     C.byref(sc), _lib.ptr(out["acc_map"]), _lib.ptr(ws), ws.numel(),
                                                     _stream(dev)), "tvm_shade_fwd_scatter")
            else:
                head = self.packed_ref_head() if self.ref_kernel else None
                if head is None:
                    raise _lib.TvmError("peer placement needs a fused shading kernel (MLP_Fea, or Ref with app_dim 27)")
                _lib.check(lib.tvm_shade_ref_fwd_scatter(C.byref(d), C.byref(head[0]), _lib.ptr(rays), n, rays.shape[1],
                                                         _lib.ptr(bg), C.byref(sc), _lib.ptr(out["acc_map"]), _lib.ptr(ws),
                                                         ws.numel(), _stream(dev)), "tvm_shade_ref_fwd_scatter")
            if keep_workspace:
                out["workspace"] = self.workspace_views(d, ws, n)
            return out
        if keep_workspace or not self.native_shade:
            views = self.workspace_views(d, ws, n)
            if keep_workspace:
                out["workspace"] = views
            if not self.native_shade and n > 0:
                head = self.packed_ref_head() if self.ref_kernel else None
                if head is not None:        # fused `Ref` tail (csrc/shade_ref.cu)
                    _lib.check(lib.tvm_shade_ref_fwd(C.byref(d), C.byref(head[0]), _lib.ptr(rays), n, rays.shape[1],
                                                     _lib.ptr(bg), _lib.ptr(out["rgb_map"]), _lib.ptr(out["depth_map"]),
                                                     _lib.ptr(out["acc_map"]), _lib.ptr(ws), ws.numel(), _stream(dev)),
                               "tvm_shade_ref_fwd")
                elif self.ref_kernel:
                    raise _lib.TvmError("this `Ref` head configuration has no fused tail kernel (it needs predicted normals "
                                        "and app_dim = 27) and the render path has no eager fallback; ref_kernel = False "
                                        "runs the torch-op tail explicitly (cross-checks only)")
                else:           # explicit cross-check switch (tests, scripts/bench_ref_head.py)
                    self._torch_shade_tail(rays, views, bg, out)
        return out

    def _torch_shade_tail(self, rays, views, bg, out):
        """Per-ray tail of TensorBase.forward (tensorBase.py:886-908) for heads without a fused kernel (`Ref`)."""
        feat = F.linear(views["ray_feat"], self.basis_mat.weight)
        rgb, _ = self.renderModule(None, rays[:, 3:6], feat, None)
        rgb = rgb * (views["app_count"] > 0).to(rgb.dtype)[:, None]
        acc = views["acc"]
        out["rgb_map"].copy_((rgb * acc[..., None] + bg * (1.0 - acc[..., None])).clamp(0, 1))
        out["depth_map"].copy_(views["depth"] + (1.0 - acc) * rays[..., -1])
        out["acc_map"].copy_(acc)

    def workspace_views(self, d, ws, n):
        """Typed views of the march-stage outputs inside a workspace buffer."""
        lib = _lib.load()
        offs = [C.c_size_t(0) for _ in range(6)]
        _lib.check(lib.tvm_workspace_layout(C.byref(d), n, *[C.byref(o) for o in offs]), "tvm_workspace_layout")
        ta = sum(self.app_n_comp)
        o = [x.value for x in offs]

        def view(off, count, dtype):
            return ws[off:off + count * 4].view(dtype)
        return {"ray_feat": view(o[0], n * ta, torch.float32).view(n, ta), "acc": view(o[1], n, torch.float32),
                "depth": view(o[2], n, torch.float32), "sigma_count": view(o[3], n, torch.int32),
                "app_count": view(o[4], n, torch.int32), "occ_count": view(o[5], n, torch.int32), "buffer": ws}

    # ------------------------------------------------------------------ the reference's forward
    def forward(self, rays_chunk, white_bg=False, bg_color=None, is_train=False, ndc_ray=False, sample_func=None,
                N_samples=-1, jitter=None):
        """TensorBase.forward (models/tensorBase.py:775-917): returns
        (rgb_map, depth_map, acc_map, alpha, z_vals, dists).

        `jitter` ([N] or [N,1], U[0,1)) is an extension: when `is_train` and it is None it is drawn with
        torch.rand on the rays' device, one value per ray, as the reference does (:507-509)."""
        if ndc_ray:
            raise NotImplementedError("ndc_ray sampling is outside the B200 render path (LLFF only, SURVEY.md 2.1)")
        point_samples = False
        if sample_func is not None:
            if getattr(sample_func, "__func__", None) is not FieldOpsMixin.sample_point_color or \
                    getattr(sample_func, "__self__", None) is not self:
                raise NotImplementedError("only sample_func=model.sample_point_color is on the B200 render path")
            point_samples = True                       # sample_point_color (tensorBase.py:623-638): no jitter
            is_train = False
            if N_samples <= 0:
                N_samples = 20
        if is_train and jitter is None:
            jitter = torch.rand(rays_chunk.shape[0], device=rays_chunk.device)
        if not is_train:
            jitter = None
        needs_grad = torch.is_grad_enabled() and (
            rays_chunk.requires_grad or any(p.requires_grad for p in self.parameters()))
        want_samples = is_train or point_samples or bool(self.eval_sample_outputs)
        if needs_grad:
            from .autograd import render_with_grad, render_with_grad_torch_tail
            if self.native_shade:
                out = render_with_grad(self, rays_chunk, white_bg, bg_color, N_samples, jitter, point_samples,
                                       want_samples=want_samples)
            else:
                out = render_with_grad_torch_tail(self, rays_chunk, white_bg, bg_color, N_samples, jitter, point_samples,
                                                  want_samples=want_samples)
        else:
            o = self.render_eval(rays_chunk, N_samples=N_samples, white_bg=white_bg, bg_color=bg_color, jitter=jitter,
                                 sample_outputs=want_samples, point_samples=point_samples)
            out = (o["rgb_map"], o["depth_map"], o["acc_map"], o.get("alpha"), o.get("z_vals"), o.get("dists"))
        if point_samples:
            # the reference's sampler returns ONE broadcast row of z values ([1,S], tensorBase.py:628-638)
            out = out[:4] + (out[4][:1], out[5][:1])
        return out
