"""`Ref` shading head (Ref-NeRF style: predicted normals, tint / diffuse / roughness heads, integrated directional
encoding, sRGB tone map) — what configs/lego.txt:25 and truck.txt:26 select and what the IFFNeRF pipeline asserts
(models/ref.py:48-155, models/ref_utils.py:6-112, models/image.py:6-13).  SURVEY.md §8f row 1.

Same constructor, parameter names (`state_dict` keys) and RNG draw order as the reference.  This fork evaluates the
head ONCE PER RAY on the accumulated feature (models/tensorBase.py:886-896), i.e. ~4 kFLOP/ray: with this head the
march stage (>90 % of the work) runs on the CUDA kernels and the head itself as torch ops on the same device.
"""
from __future__ import annotations

import math

import torch


class _Shift(torch.nn.Module):
    def __init__(self, value: float):
        super().__init__()
        self.value = value

    def forward(self, x):
        return x + self.value


class _Scale(torch.nn.Module):
    def __init__(self, value: float):
        super().__init__()
        self.value = value

    def forward(self, x):
        return x * self.value


class _UnitNorm(torch.nn.Module):
    def forward(self, x):
        return torch.nn.functional.normalize(x, p=2, dim=-1)


def reflect(viewdirs, normals):
    """u = 2 (n.v) n - v   (models/ref_utils.py:6-19)."""
    dot = torch.bmm(normals.view(-1, 1, 3), viewdirs.view(-1, 3, 1))[..., 0]
    return torch.multiply(2.0 * dot, normals) - viewdirs


def linear_to_srgb(linear, eps=None):
    """models/image.py:6-13."""
    if eps is None:
        eps = torch.finfo(linear.dtype).eps
    low = 323 / 25 * linear
    high = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, low, high)


class IntegratedDirEnc(torch.nn.Module):
    """Integrated directional encoding (models/ref_utils.py:22-112): spherical harmonics of degrees 2^i attenuated
    by exp(-l(l+1)/2 * roughness).  `ml_array` / `mat` are (frozen) Parameters, as in the reference state_dict."""

    @staticmethod
    def _binom(a, k):
        return torch.prod(a - torch.arange(k)) / math.factorial(k)

    @staticmethod
    def _legendre(l, m, k):
        return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
                * IntegratedDirEnc._binom(0.5 * (l + k + m - 1.0), l))

    @staticmethod
    def _sph(l, m, k):
        return (math.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m)))
                * IntegratedDirEnc._legendre(l, m, k))

    def __init__(self, deg_view: int):
        super().__init__()
        pairs = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
        self.ml_array = torch.nn.Parameter(torch.tensor(pairs).T, requires_grad=False)
        l_max = 2 ** (deg_view - 1)
        mat = torch.zeros((l_max + 1, len(pairs)))
        for col, (m, l) in enumerate(self.ml_array.T):
            for k in range(l - m + 1):
                mat[k, col] = IntegratedDirEnc._sph(l, m, k)
        self.mat = torch.nn.Parameter(mat, requires_grad=False)

    def forward(self, xyz, kappa_inv):
        x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
        z_pows = torch.pow(z, torch.arange(self.mat.shape[0], dtype=z.dtype, device=z.device)[None, :])
        xy_pows = torch.pow((x + 1j * y), self.ml_array[0, :])
        harmonics = xy_pows * torch.matmul(z_pows, self.mat)
        sigma = 0.5 * self.ml_array[1, :] * (self.ml_array[1, :] + 1)
        return torch.view_as_real(harmonics * torch.exp(-sigma * kappa_inv))


class Ref(torch.nn.Module):
    """models/ref.py:48-155."""

    def __init__(self, in_channels, viewpe=6, feature_c=128, deg_view=4, predicted_normals=True,
                 rgb_premultiplier=1.0, rgb_bias=0.0):
        super().__init__()
        self.dir_enc_fn = IntegratedDirEnc(deg_view)
        self.rgb_padding = 0.001
        self.in_mlpC = (3 + 2 * viewpe * 3) + in_channels
        self.viewpe = viewpe
        lin = torch.nn.Linear
        self.diffuse_color_mlp = torch.nn.Sequential(lin(in_channels, 3), _Shift(-math.log(3.0)), torch.nn.Sigmoid())
        self.tint_color_mlp = torch.nn.Sequential(lin(in_channels, 3), torch.nn.Sigmoid())
        self.roughness_mlp = torch.nn.Sequential(lin(in_channels, 1), _Shift(-1.0), torch.nn.Softplus())
        self.bottleneck_mlp = lin(in_channels, feature_c)
        self.predicted_normals = predicted_normals
        if predicted_normals:
            self.normal_mlp = torch.nn.Sequential(lin(in_channels, 3), _UnitNorm(), _Scale(-1))
        n_dir = sum((2 ** i) + 1 for i in range(deg_view)) * 2 + 1
        spec = [lin(feature_c + n_dir, 3)]
        if rgb_premultiplier < 1.0 - 1e-7 or rgb_premultiplier > 1.0 + 1e-7:
            spec.append(_Scale(rgb_premultiplier))
        if rgb_bias > 1e-7:
            spec.append(_Shift(rgb_bias))
        spec.append(torch.nn.Sigmoid())
        self.specular_mlp = torch.nn.Sequential(*spec)

    def forward(self, pts, viewdirs, features, normals):
        if normals is None and self.predicted_normals:
            normals = self.normal_mlp(features)
        tint = self.tint_color_mlp(features)
        roughness = self.roughness_mlp(features)
        bottleneck = self.bottleneck_mlp(features)
        refdirs = reflect(-viewdirs, normals)
        enc = self.dir_enc_fn(refdirs, roughness)
        n_dot_v = torch.bmm(normals.view(-1, 1, 3), viewdirs.view(-1, 3, 1))[..., 0]
        x = torch.cat([bottleneck, enc.view(enc.shape[0], math.prod(enc.shape[1:])), n_dot_v], dim=-1)
        specular = tint * self.specular_mlp(x)
        diffuse = self.diffuse_color_mlp(features)
        rgb = torch.clip(linear_to_srgb(specular + diffuse), 0.0, 1.0)
        return rgb * (1 + 2 * self.rgb_padding) - self.rgb_padding, None

    def compute_normals(self, features):
        return -self.normal_mlp(features)
