"""CPU checks of the kernels' per-sample math (the SAME .cuh source, built for the host) against the
reference-generated goldens and the oracle.  The CUDA path itself is tested in test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostcheck") / "libhostcheck.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                    os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")], check=True)
    return C.CDLL(out)


def _host_desc(hc, fld):
    m = H.module_from_field(fld, "cpu")
    d, buf = H.host_pack_factors(m)
    keep = [buf]
    d.factors = buf.ctypes.data
    if fld.occupancy is not None:
        vol = np.ascontiguousarray(fld.occupancy.volume.numpy(), dtype=np.float32)
        dz, dy, dx = vol.shape
        cells = np.zeros(vol.shape, dtype=np.uint8)
        hc.hc_pack_occupancy(C.c_void_p(vol.ctypes.data), dx, dy, dz, C.c_void_p(cells.ctypes.data))
        d.occ_cells = cells.ctypes.data
        d.occ_dims[:] = [dx, dy, dz]
        d.occ_lo[:] = m.alphaMask._lo
        d.occ_inv[:] = m.alphaMask._inv
        keep.append(cells)
    return m, d, keep


def _mask(hc, d, rays, S, jitter=None):
    rays = np.ascontiguousarray(rays.numpy(), dtype=np.float32)
    n = rays.shape[0]
    bits = np.zeros((n, (S + 31) // 32), dtype=np.uint32)
    counts = np.zeros(n, dtype=np.int32)
    jp = None if jitter is None else C.c_void_p(np.ascontiguousarray(jitter, dtype=np.float32).ctypes.data)
    hc.hc_sample_mask(C.byref(d), C.c_void_p(rays.ctypes.data), C.c_longlong(n), rays.shape[1], S, jp,
                      C.c_void_p(bits.ctypes.data), C.c_void_p(counts.ctypes.data))
    return bits, counts


CASES = {
    "c1_dense_mask": lambda: fx.config1(0.0, "sphere", 6),
    "c1_refdefault_nomask": lambda: fx.config1(-10.0, None, 6),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_sample_mask_bit_exact_vs_reference_golden(hc, name):
    fld, rays = CASES[name]()
    g = H.golden(name)
    H.check_params(fld, g)
    m, d, keep = _host_desc(hc, fld)
    assert m.nSamples == int(g["n_samples"])
    assert np.float32(m._host["step"]) == g["step_size"]
    bits, counts = _mask(hc, d, rays, m.nSamples)
    assert np.array_equal(bits, g["valid_bits"])          # bit-exact, every sample of every ray
    assert np.array_equal(counts, g["valid_count"])


def test_sample_mask_bit_exact_noncubic_and_jitter(hc):
    """Non-cubic grid + occupancy with its own dims (config 4) and the train-mode jitter, vs the oracle."""
    fld, rays = fx.config4(H=54, W=96)
    m, d, keep = _host_desc(hc, fld)
    torch.manual_seed(5)
    jit = torch.rand(rays.shape[0], 1)
    for jitter in (None, jit):
        _, _, valid = orc.sample_along_rays(fld, rays[:, :3], rays[:, 3:6], m.nSamples, jitter)
        pts, _, _ = orc.sample_along_rays(fld, rays[:, :3], rays[:, 3:6], m.nSamples, jitter)
        keepm = torch.zeros_like(valid)
        keepm[valid] = orc.occupancy_value(fld.occupancy, pts[valid]) > 0
        bits, counts = _mask(hc, d, rays, m.nSamples, None if jitter is None else jitter.numpy().reshape(-1))
        assert np.array_equal(H.unpack_bits(bits, m.nSamples), keepm.numpy())
        assert counts.sum() > 0


def test_march_math_matches_reference_golden(hc):
    """Host walk of the march recurrence (same gather/tap source as the kernel) + oracle shading == golden rgb."""
    fld, rays = fx.config1(0.0, "sphere", 6)
    g = H.golden("c1_dense_mask")
    m, d, keep = _host_desc(hc, fld)
    sel = np.arange(0, rays.shape[0], 7)[:600]
    r = np.ascontiguousarray(rays.numpy()[sel], dtype=np.float32)
    n, S = r.shape[0], m.nSamples
    feat = np.zeros((n, 144), dtype=np.float32)
    acc = np.zeros(n, dtype=np.float32)
    dep = np.zeros(n, dtype=np.float32)
    napp = np.zeros(n, dtype=np.int32)
    hc.hc_march(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(n), 6, S, None, C.c_void_p(feat.ctypes.data),
                C.c_void_p(acc.ctypes.data), C.c_void_p(dep.ctypes.data), None, C.c_void_p(napp.ctypes.data))
    np.testing.assert_allclose(acc, g["acc_map"][sel], atol=2e-5)
    # app_mask is not bit-reproducible, but counts must agree to within a few borderline samples
    assert np.abs(napp - g["app_count"][sel]).max() <= 2
    f27 = torch.from_numpy(feat) @ fld.basis.T
    lit = torch.from_numpy(napp > 0)
    rgb = torch.zeros(n, 3)
    rgb[lit] = orc.shade(fld, torch.from_numpy(r[:, 3:6])[lit], f27[lit])
    a = torch.from_numpy(acc)[:, None]
    rgb = (rgb * a + 1.0 * (1 - a)).clamp(0, 1)
    np.testing.assert_allclose(rgb.numpy(), g["rgb_map"][sel], atol=1e-4)
    depth = dep + (1 - acc) * r[:, -1]
    np.testing.assert_allclose(depth, g["depth_map"][sel], atol=1e-4)
