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
        cdims = [(dx + 15) // 16, (dy + 15) // 16, (dz + 15) // 16]
        coarse = np.zeros(cdims[::-1], dtype=np.uint8)
        hc.hc_pack_coarse(C.c_void_p(cells.ctypes.data), dx, dy, dz, C.c_void_p(coarse.ctypes.data))
        d.occ_cells = cells.ctypes.data
        d.occ_dims[:] = [dx, dy, dz]
        d.occ_lo[:] = m.alphaMask._lo
        d.occ_inv[:] = m.alphaMask._inv
        d.occ_coarse = coarse.ctypes.data
        d.occ_cdims[:] = cdims
        keep += [cells, coarse]
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


@pytest.mark.parametrize("case", ["c1_mask", "c1_nomask", "c4_jitter"])
def test_block_skip_flags_are_conservative(hc, case):
    """The empty-space skip test never discards a 32-sample block that holds a valid sample (so skipping cannot
    change ray_valid), and it does discard most of the empty ones."""
    if case == "c1_mask":
        fld, rays = fx.config1(0.0, "sphere", 6)
    elif case == "c1_nomask":
        fld, rays = fx.config1(-10.0, None, 6)
    else:
        fld, rays = fx.config4(H=54, W=96)
    rays = rays[::3].contiguous()
    m, d, keep = _host_desc(hc, fld)
    S = m.nSamples
    jit = None
    if case == "c4_jitter":
        torch.manual_seed(9)
        jit = np.ascontiguousarray(torch.rand(rays.shape[0]).numpy(), dtype=np.float32)
    bits, counts = _mask(hc, d, rays, S, jit)
    r = np.ascontiguousarray(rays.numpy(), dtype=np.float32)
    nblk = (S + 31) // 32
    flags = np.zeros((r.shape[0], nblk), dtype=np.uint8)
    hc.hc_block_flags(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(r.shape[0]), r.shape[1], S,
                      None if jit is None else C.c_void_p(jit.ctypes.data), C.c_void_p(flags.ctypes.data))
    has_valid = bits != 0                                   # [n, nblk]: block holds at least one valid sample
    assert not np.any(has_valid & (flags == 0))             # conservative: no valid block is ever skipped
    assert has_valid.sum() > 0
    if fld.occupancy is not None:
        assert flags.sum() <= 3.0 * has_valid.sum() + 2 * r.shape[0]   # ...and the filter is reasonably tight


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


def test_run_cache_appearance_matches_per_sample_path(hc):
    """app_gather_kernel's gathers (quad runs over the ray's sample list with a register texel cache,
    tvm_gather.cuh::app_run_plane) emulated lane by lane on the host == the per-sample accumulation of the fused
    kernel (app_accumulate_taps), on cubic and non-cubic grids."""
    for fld, rays in (fx.config1(0.0, "sphere", 6), fx.config4(24, 40)):
        m, d, keep = _host_desc(hc, fld)
        sel = np.arange(0, rays.shape[0], 5)[:400]
        r = np.ascontiguousarray(rays.numpy()[sel], dtype=np.float32)
        n, S = r.shape[0], m.nSamples
        outs = []
        for mode in (0, 1):
            hc.hc_set_app_octets(mode)
            feat = np.zeros((n, 144), dtype=np.float32)
            acc = np.zeros(n, dtype=np.float32)
            dep = np.zeros(n, dtype=np.float32)
            napp = np.zeros(n, dtype=np.int32)
            hc.hc_march(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(n), r.shape[1], S, None,
                        C.c_void_p(feat.ctypes.data), C.c_void_p(acc.ctypes.data), C.c_void_p(dep.ctypes.data), None,
                        C.c_void_p(napp.ctypes.data))
            outs.append((feat, acc, napp))
        hc.hc_set_app_octets(0)
        assert outs[0][2].sum() > 1000 and np.array_equal(outs[0][2], outs[1][2])
        scale = np.abs(outs[0][0]).max()
        np.testing.assert_allclose(outs[1][0], outs[0][0], atol=2e-6 * max(scale, 1.0))


def _unpack_host_grads(m, d, gbuf):
    out = {}
    for k in range(3):
        for name, off, t in ((f"density_plane.{k}", d.dplane_off[k], m.density_plane[k]),
                             (f"density_line.{k}", d.dline_off[k], m.density_line[k]),
                             (f"app_plane.{k}", d.aplane_off[k], m.app_plane[k]),
                             (f"app_line.{k}", d.aline_off[k], m.app_line[k])):
            _, C_, H_, W_ = t.shape
            pitch = (W_ | 1) if W_ > 1 else 1            # plane rows are padded to an odd pitch (tvm_plane_pitch)
            out[name] = gbuf[off:off + C_ * H_ * pitch].reshape(H_, pitch, C_)[:, :W_].transpose(2, 0, 1)[None]
    return out


def _bwd_setup(hc, cols=6, n=160, jitter=False, seed=3):
    fld, rays = fx.config1(0.0, "sphere", cols)
    sel = torch.arange(1500, 1500 + 37 * n, 37)
    rays = rays[sel].contiguous()
    m, d, keep = _host_desc(hc, fld)
    torch.manual_seed(seed)
    jit = torch.rand(n, 1) if jitter else None
    target = torch.rand(n, 3)
    return fld, rays, m, d, keep, jit, target


def _host_forward(hc, d, r, S, jit):
    n = r.shape[0]
    feat = np.zeros((n, 144), dtype=np.float32)
    acc = np.zeros(n, dtype=np.float32)
    dep = np.zeros(n, dtype=np.float32)
    alpha = np.zeros((n, S), dtype=np.float32)
    jp = None if jit is None else C.c_void_p(jit.ctypes.data)
    hc.hc_march(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(n), r.shape[1], S, jp,
                C.c_void_p(feat.ctypes.data), C.c_void_p(acc.ctypes.data), C.c_void_p(dep.ctypes.data),
                C.c_void_p(alpha.ctypes.data), None)
    return feat, acc, alpha


def test_march_backward_factor_grads_match_oracle_autograd(hc):
    """train.py-style loss (MSE + 0.1*mean(exp|alpha|)) on jittered rays: the host walk of the backward recurrence
    (same scatter source as the CUDA kernel) reproduces the oracle's autograd factor gradients."""
    fld, rays, m, d, keep, jit, target = _bwd_setup(hc, jitter=True)
    S = m.nSamples + 3
    factors = fld.density_plane + fld.density_line + fld.app_plane + fld.app_line
    for p in factors:
        p.requires_grad_(True)
    o = orc.render_chunk(fld, rays, bg_color=torch.ones(3), n_samples=S, jitter=jit)
    o["ray_feat"].retain_grad()
    o["acc_map"].retain_grad()
    loss = orc.train_loss(o, target)
    loss.backward()                       # (one pass only: retain_grad hooks accumulate on every pass)
    g27, g_acc = o["ray_feat"].grad, o["acc_map"].grad
    grads = [p.grad.clone() for p in factors]
    for p in factors:
        p.requires_grad_(False)
        p.grad = None
    r = np.ascontiguousarray(rays.numpy(), dtype=np.float32)
    jn = np.ascontiguousarray(jit.numpy().reshape(-1), dtype=np.float32)
    feat, acc, alpha = _host_forward(hc, d, r, S, jn)
    np.testing.assert_allclose(alpha, o["alpha"].detach().numpy(), atol=1e-5)
    gF = np.ascontiguousarray((g27 @ fld.basis).numpy(), dtype=np.float32)
    ga = np.ascontiguousarray(g_acc.numpy(), dtype=np.float32)
    dal = np.ascontiguousarray(0.1 / alpha.size * np.exp(np.abs(alpha)) * np.sign(alpha), dtype=np.float32)
    gbuf = np.zeros(int(d.n_factor_floats), dtype=np.float32)
    hc.hc_march_bwd(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(r.shape[0]), r.shape[1], S,
                    C.c_void_p(jn.ctypes.data), C.c_void_p(feat.ctypes.data), C.c_void_p(acc.ctypes.data),
                    C.c_void_p(gF.ctypes.data), C.c_void_p(ga.ctypes.data), C.c_void_p(dal.ctypes.data),
                    C.c_void_p(gbuf.ctypes.data), None)
    mine = _unpack_host_grads(m, d, gbuf)
    names = ([f"density_plane.{k}" for k in range(3)] + [f"density_line.{k}" for k in range(3)]
             + [f"app_plane.{k}" for k in range(3)] + [f"app_line.{k}" for k in range(3)])
    for nme, g in zip(names, grads):
        ref = g.numpy()
        scale = np.abs(ref).max()
        assert scale > 0, nme
        err = np.abs(mine[nme] - ref).max()
        assert err <= 2e-3 * scale, (nme, err, scale)


def test_march_backward_ray_grads_match_oracle_autograd(hc):
    """Pose mode (inerf/estimate_pose_inerf.py:164-178): frozen factors, MSE loss, gradient w.r.t. the rays."""
    fld, rays, m, d, keep, _, target = _bwd_setup(hc, cols=7, n=120)
    S = m.nSamples
    bg = torch.tensor([0.2, 0.5, 0.9])
    rr = rays.clone().requires_grad_(True)
    o = orc.render_chunk(fld, rr, bg_color=bg)
    o["ray_feat"].retain_grad()
    o["acc_map"].retain_grad()
    loss = torch.mean((o["rgb_map"] - target) ** 2)
    loss.backward()
    g27, g_acc = o["ray_feat"].grad, o["acc_map"].grad
    # the shading head's own contribution to d(viewdirs): autograd on the per-ray stage alone
    view = rays[:, 3:6].clone().requires_grad_(True)
    lit = o["app_mask"].any(-1)
    rgb = torch.zeros(rays.shape[0], 3)
    rgb[lit] = orc.shade(fld, view[lit], o["ray_feat"].detach()[lit])
    a = o["acc_map"].detach()[:, None]
    l2 = torch.mean(((rgb * a + bg * (1 - a)).clamp(0, 1) - target) ** 2)
    g_view, = torch.autograd.grad(l2, view)
    r = np.ascontiguousarray(rays.numpy(), dtype=np.float32)
    feat, acc, alpha = _host_forward(hc, d, r, S, None)
    gF = np.ascontiguousarray((g27 @ fld.basis).numpy(), dtype=np.float32)
    ga = np.ascontiguousarray(g_acc.numpy(), dtype=np.float32)
    g_rays = np.zeros((r.shape[0], 6), dtype=np.float32)
    hc.hc_march_bwd(C.byref(d), C.c_void_p(r.ctypes.data), C.c_longlong(r.shape[0]), r.shape[1], S, None,
                    C.c_void_p(feat.ctypes.data), C.c_void_p(acc.ctypes.data), C.c_void_p(gF.ctypes.data),
                    C.c_void_p(ga.ctypes.data), None, None, C.c_void_p(g_rays.ctypes.data))
    g_rays[:, 3:6] += g_view.numpy()
    ref = rr.grad.numpy()[:, :6]
    scale = np.abs(ref).max()
    assert scale > 0
    assert np.abs(g_rays - ref).max() <= 5e-3 * scale, (np.abs(g_rays - ref).max(), scale)


def test_point_sample_mode_mask_bit_exact(hc):
    """sample_point_color sampler (TVM_F_POINT_SAMPLES): 20 samples centred on the origin, vs the oracle."""
    from oracle.make_golden import point_rays
    fld, _ = fx.config1(0.0, "sphere", 6)
    rays = point_rays(fld, 2000)
    m, d, keep = _host_desc(hc, fld)
    pts, z, valid = orc.sample_around_points(fld, rays[:, :3], rays[:, 3:6], 20)
    occ = torch.zeros_like(valid)
    occ[valid] = orc.occupancy_value(fld.occupancy, pts[valid]) > 0
    hc.hc_set_point_samples(1)
    try:
        bits, counts = _mask(hc, d, rays, 20)
    finally:
        hc.hc_set_point_samples(0)
    assert np.array_equal(H.unpack_bits(bits, 20), occ.numpy())
    assert counts.sum() > 0
