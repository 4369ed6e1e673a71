"""Parameter container of the `Ref` shading head (what configs/lego.txt:25 and truck.txt:26 select and the IFFNeRF
pipeline asserts; models/ref.py:48-155, models/ref_utils.py:6-112, models/image.py:6-13).

The head is evaluated by the fused kernels of csrc/shade_ref.cu in both directions.  This module only
  * owns the parameters under the reference's `state_dict` keys (`renderModule.<name>_mlp.0.{weight,bias}`,
    `bottleneck_mlp.*`, `dir_enc_fn.{ml_array,mat}`) — created in the reference's order, so one seed gives the reference's
    weights —, the directional-encoding tables and the head's scalar constants the kernels are fed with, and
  * offers `forward` as a plain-tensor evaluation of the same function, used as the independent cross-check of the
    kernels (tests, `tensorf.ref_kernel = False`) and `compute_normals` for the pose pipeline
    (pose_estimation/sampling.py:535-541).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ide_tables(deg_view: int):
    """(m, l) pairs and the z-polynomial coefficient table of the integrated directional encoding: degrees l = 2^i,
    orders m = 0..l; column (m, l) holds the coefficients c_k of z^k in the real polynomial factor of Y_l^m,
    c_k = sqrt((2l+1)(l-m)! / (4 pi (l+m)!)) * (-1)^m 2^l l!/(k!(l-k-m)!) * binom((l+k+m-1)/2, l)."""
    pairs = [(m, 2 ** i) for i in range(deg_view) for m in range(2 ** i + 1)]
    l_max = 2 ** (deg_view - 1)
    table = torch.zeros((l_max + 1, len(pairs)))
    f32 = lambda v: torch.tensor(v, dtype=torch.float32)
    # The table is a checkpointed tensor, so it is built in the reference's rounding sequence (integer products exact
    # in int64, every division and the later products in fp32, the square root in double on the fp32 radicand).
    for col, (m, l) in enumerate(pairs):
        radicand = f32(2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * math.pi * math.factorial(l + m))
        norm = math.sqrt(float(radicand))
        lead = torch.tensor((-1) ** m * 2 ** l * math.factorial(l), dtype=torch.int64)
        for k in range(l - m + 1):
            top = 0.5 * (l + k + m - 1.0)
            gen_binom = torch.prod(top - torch.arange(l)) / math.factorial(l)          # binom(top, l) for real `top`
            table[k, col] = norm * (lead / math.factorial(k) / math.factorial(l - k - m) * gen_binom)
    return torch.tensor(pairs).T, table


class IntegratedDirEnc(torch.nn.Module):
    """Holder of the (frozen) encoding tables under the reference's parameter names."""

    def __init__(self, deg_view: int):
        super().__init__()
        ml, table = _ide_tables(deg_view)
        self.ml_array = torch.nn.Parameter(ml, requires_grad=False)
        self.mat = torch.nn.Parameter(table, requires_grad=False)

    def forward(self, direction, roughness):
        """[N,3] unit directions, [N,1] roughness -> [N, pairs, 2] (real, imaginary), attenuated by exp(-l(l+1)/2 r).
        Real arithmetic throughout: z^k by running products, (x + iy)^m by the complex-multiplication recurrence —
        the formulation the kernel uses."""
        x, y, z = direction[:, 0:1], direction[:, 1:2], direction[:, 2:3]
        n_k = self.mat.shape[0]
        z_pow = torch.cat([torch.ones_like(z), z.expand(-1, n_k - 1)], dim=1).cumprod(dim=1)        # [N, l_max+1]
        poly = z_pow @ self.mat                                                                   # [N, pairs]
        m_max = int(self.ml_array[0].max())
        re, im = [torch.ones_like(x)], [torch.zeros_like(x)]
        for _ in range(m_max):
            re, im = re + [re[-1] * x - im[-1] * y], im + [re[-1] * y + im[-1] * x]
        re, im = torch.cat(re, dim=1), torch.cat(im, dim=1)                                       # [N, m_max+1]
        m_idx, l = self.ml_array[0].long(), self.ml_array[1].to(direction.dtype)
        damp = torch.exp(-0.5 * l * (l + 1.0) * roughness)                                        # [N, pairs]
        return torch.stack([re[:, m_idx] * poly * damp, im[:, m_idx] * poly * damp], dim=-1)


class Ref(torch.nn.Module):
    """Ref-NeRF style head: predicted normal, tint / diffuse / roughness, a bottleneck, a specular layer on
    [bottleneck | directional encoding of the reflected view | n.v], sRGB tone map, rgb padding."""

    diffuse_shift = -math.log(3.0)        # bias inside the diffuse sigmoid  (models/ref.py:66-70)
    rough_shift = -1.0                    # bias inside the roughness softplus

    def __init__(self, in_channels, viewpe=6, feature_c=128, deg_view=4, predicted_normals=True,
                 rgb_premultiplier=1.0, rgb_bias=0.0):
        super().__init__()
        self.dir_enc_fn = IntegratedDirEnc(deg_view)
        self.rgb_padding = 0.001
        self.rgb_premultiplier, self.rgb_bias = float(rgb_premultiplier), float(rgb_bias)
        self.in_mlpC = (3 + 2 * viewpe * 3) + in_channels
        self.viewpe = viewpe
        self.predicted_normals = predicted_normals

        def head(n_out):           # `<name>.0.{weight,bias}` like the reference's Sequential(Linear, activation...)
            return torch.nn.Sequential(torch.nn.Linear(in_channels, n_out))
        # creation order = the reference's (RNG stream)
        self.diffuse_color_mlp = head(3)
        self.tint_color_mlp = head(3)
        self.roughness_mlp = head(1)
        self.bottleneck_mlp = torch.nn.Linear(in_channels, feature_c)
        if predicted_normals:
            self.normal_mlp = head(3)
        n_dir = 2 * int(self.dir_enc_fn.ml_array.shape[1]) + 1
        self.specular_mlp = torch.nn.Sequential(torch.nn.Linear(feature_c + n_dir, 3))

    # ---- the same function as plain tensor ops (cross-check of csrc/shade_ref.cu)
    def _normal(self, features):
        return -F.normalize(self.normal_mlp[0](features), p=2, dim=-1)

    def compute_normals(self, features):
        """models/ref.py: the outward normal (the head uses its negation internally)."""
        return -self._normal(features)

    def forward(self, pts, viewdirs, features, normals):
        if normals is None and self.predicted_normals:
            normals = self._normal(features)
        tint = torch.sigmoid(self.tint_color_mlp[0](features))
        roughness = F.softplus(self.roughness_mlp[0](features) + self.rough_shift)
        diffuse = torch.sigmoid(self.diffuse_color_mlp[0](features) + self.diffuse_shift)
        n_dot_v = (normals * viewdirs).sum(-1, keepdim=True)
        reflected = viewdirs - 2.0 * n_dot_v * normals            # reflection of -v about n, written out
        enc = self.dir_enc_fn(reflected, roughness)
        spec_in = torch.cat([self.bottleneck_mlp(features), enc.reshape(enc.shape[0], -1), n_dot_v], dim=-1)
        spec = self.specular_mlp[0](spec_in)
        if abs(self.rgb_premultiplier - 1.0) > 1e-7:
            spec = spec * self.rgb_premultiplier
        if self.rgb_bias > 1e-7:
            spec = spec + self.rgb_bias
        linear = tint * torch.sigmoid(spec) + diffuse
        eps = torch.finfo(linear.dtype).eps                       # sRGB transfer function (models/image.py:6-13)
        srgb = torch.where(linear <= 0.0031308, (323.0 / 25.0) * linear,
                           (211.0 * linear.clamp(min=eps) ** (5.0 / 12.0) - 11.0) / 200.0)
        return srgb.clamp(0.0, 1.0) * (1.0 + 2.0 * self.rgb_padding) - self.rgb_padding, None
