"""Shared test helpers: oracle fixtures -> product module, golden loading, host-side packing."""
import contextlib
import io
import os

import numpy as np
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def module_from_field(fld, device):
    """Builds the product TensorVMSplit holding exactly the oracle field's parameters."""
    import iffnerf_b200 as I
    with contextlib.redirect_stdout(io.StringIO()):
        m = I.TensorVMSplit(fld.aabb.clone().to(device), list(fld.grid), device,
                            density_n_comp=[p.shape[1] for p in fld.density_plane],
                            appearance_n_comp=[p.shape[1] for p in fld.app_plane], app_dim=fld.basis.shape[0],
                            near_far=list(fld.near_far),
                            shadingMode="MLP_Fea", alphaMask_thres=1e-4, density_shift=fld.density_shift,
                            distance_scale=fld.distance_scale, rayMarch_weight_thres=fld.weight_thres, pos_pe=6,
                            view_pe=fld.view_pe, fea_pe=fld.fea_pe, featureC=128, step_ratio=fld.step_ratio,
                            fea2denseAct=fld.fea2dense)
    sd = {}
    for k in range(3):
        sd[f"density_plane.{k}"] = fld.density_plane[k]
        sd[f"density_line.{k}"] = fld.density_line[k]
        sd[f"app_plane.{k}"] = fld.app_plane[k]
        sd[f"app_line.{k}"] = fld.app_line[k]
    sd["basis_mat.weight"] = fld.basis
    for i, li in enumerate((0, 2, 4)):
        sd[f"renderModule.mlp.{li}.weight"] = fld.mlp_w[i]
        sd[f"renderModule.mlp.{li}.bias"] = fld.mlp_b[i]
    m.load_state_dict(sd)
    if fld.occupancy is not None:
        m.alphaMask = I.AlphaGridMask(device, fld.occupancy.aabb.clone().to(device),
                                      fld.occupancy.volume.clone().to(device))
    return m


def check_params(fld, g):
    """RNG-drift guard: the regenerated fixture must have the parameters the goldens were made with."""
    np.testing.assert_allclose(fx.param_checksum(fld), g["param_checksum"], rtol=1e-10, atol=1e-10)  # double sums: thread-count dependent order


def host_pack_factors(m):
    """numpy restatement of the packed channel-last layout (include/tvm_b200.h), for host-side checks."""
    d = m._base_desc()
    buf = np.zeros(int(d.n_factor_floats), dtype=np.float32)
    for k in range(3):
        for off, t in ((d.dplane_off[k], m.density_plane[k]), (d.dline_off[k], m.density_line[k]),
                       (d.aplane_off[k], m.app_plane[k]), (d.aline_off[k], m.app_line[k])):
            a = t.detach().cpu().numpy()[0]                   # [C,H,W]
            a = a.transpose(1, 2, 0)                          # [H,W,C]
            if a.shape[1] > 1:                                # planes: rows padded to an odd pitch (tvm_plane_pitch)
                a = np.pad(a, ((0, 0), (0, (a.shape[1] | 1) - a.shape[1]), (0, 0)))
            a = np.ascontiguousarray(a).reshape(-1)
            buf[off:off + a.size] = a
    return d, buf


def unpack_bits(words, n_samples):
    """[N, W] uint32 words -> [N, S] bool."""
    w = np.ascontiguousarray(words).view(np.uint32)
    b = np.unpackbits(w.view(np.uint8).reshape(w.shape[0], -1), axis=1, bitorder="little")
    return b[:, :n_samples].astype(bool)
