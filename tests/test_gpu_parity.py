"""Parity of the CUDA path (through the C ABI) against the reference-generated goldens and the oracle.

Tolerances (BASELINE.json north_star): valid-sample mask and counts BIT-EXACT; rgb_map / depth_map within
1e-4 max-abs in fp32.  Every test here needs a B200 (`-m gpu`)."""
import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-4

EVAL_CASES = {
    "c1_dense_mask": (lambda: fx.config1(0.0, "sphere", 6), None, True),
    "c1_refdefault_nomask": (lambda: fx.config1(-10.0, None, 6), None, True),
    "c1_dense_7col_blackbg": (lambda: fx.config1(0.0, "sphere", 7), "ray_index", None),
    "c2_sub": (lambda: fx.config2(), "ray_index", True),
    "c4_sub": (lambda: fx.config4(), "ray_index", True),
}


@pytest.fixture(scope="module")
def dev(built_lib):
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    return torch.device("cuda:0")


_cache = {}


def _case(name, dev):
    if name not in _cache:
        _cache.clear()                       # one 300^3 model at a time is plenty
        build, idx_key, white = EVAL_CASES[name]
        fld, rays = build()
        g = H.golden(name)
        H.check_params(fld, g)
        if idx_key:
            rays = rays[torch.from_numpy(g[idx_key])].contiguous()
        _cache[name] = (fld, rays, g, white, H.module_from_field(fld, dev))
    return _cache[name]


@pytest.mark.parametrize("name", sorted(EVAL_CASES))
def test_valid_mask_and_counts_bit_exact(name, dev):
    fld, rays, g, white, m = _case(name, dev)
    bits, counts = m.sample_mask(rays.to(dev))
    assert m.nSamples == int(g["n_samples"])
    assert np.array_equal(bits.cpu().numpy().view(np.uint32), g["valid_bits"])
    assert np.array_equal(counts.cpu().numpy(), g["valid_count"])


@pytest.mark.parametrize("mlp", ["auto", "fp32"])
@pytest.mark.parametrize("early_term", [False, True])
@pytest.mark.parametrize("name", sorted(EVAL_CASES))
def test_render_matches_reference_golden(name, early_term, mlp, dev):
    """mlp='auto' is the default shading kernel (tensor cores, bf16x3 split operands); 'fp32' the SIMT FFMA kernel."""
    fld, rays, g, white, m = _case(name, dev)
    m.mlp_precision = mlp
    try:
        assert m._shade_mode() == ("tc3" if mlp == "auto" else "fp32")
        o = m.render_eval(rays.to(dev), white_bg=bool(white), early_term=early_term, want_counts=True)
        torch.cuda.synchronize()
    finally:
        m.mlp_precision = type(m).mlp_precision
    assert np.abs(o["rgb_map"].cpu().numpy() - g["rgb_map"]).max() <= TOL
    assert np.abs(o["depth_map"].cpu().numpy() - g["depth_map"]).max() <= TOL
    assert np.abs(o["acc_map"].cpu().numpy() - g["acc_map"]).max() <= TOL
    # the full valid count is bit-exact even when rays terminate early
    assert np.array_equal(o["valid_count"].cpu().numpy(), g["valid_count"])
    # app_mask depends on fp32 reduction order: borderline samples only
    assert np.abs(o["app_count"].cpu().numpy() - g["app_count"]).max() <= 2


@pytest.mark.parametrize("name", ["c1_dense_mask", "c2_sub", "c4_sub"])
def test_split_march_equals_fused_march(name, dev):
    """The default eval path (sigma-march that emits per-ray appearance lists + app_gather_kernel with its register
    texel cache) against the one-kernel march: identical sample decisions, ray_feat equal up to fp32 summation order."""
    fld, rays, g, white, m = _case(name, dev)
    outs = {}
    for split in (True, False):
        m.split_app = split
        try:
            o = m.render_eval(rays.to(dev), white_bg=bool(white), want_counts=True, keep_workspace=True)
            torch.cuda.synchronize()
        finally:
            m.split_app = type(m).split_app
        outs[split] = o
    a, b = outs[True], outs[False]
    assert torch.equal(a["valid_count"], b["valid_count"]) and torch.equal(a["app_count"], b["app_count"])
    assert torch.equal(a["acc_map"], b["acc_map"]) and torch.equal(a["workspace"]["depth"], b["workspace"]["depth"])
    fa, fb = a["workspace"]["ray_feat"], b["workspace"]["ray_feat"]
    assert int(a["app_count"].max()) <= 128          # no list overflow in these fixtures: the gather kernel did the work
    assert float((fa - fb).abs().max()) <= 2e-6 * max(1.0, float(fb.abs().max()))
    assert float((a["rgb_map"] - b["rgb_map"]).abs().max()) <= 2e-6
    assert np.abs(a["rgb_map"].cpu().numpy() - g["rgb_map"]).max() <= TOL


def test_split_march_list_overflow_takes_the_fused_pass(dev):
    """Rays with more than TVM_APP_CAP (128) appearance samples are re-marched by the fused kernel's overflow pass:
    a thin medium (distance_scale 5: alpha ~ 0.017 per sample, ~300 samples above the weight threshold)."""
    fld, rays = fx.config1(0.0, "sphere", 6)
    fld.distance_scale = 5.0
    sel = torch.arange(0, rays.shape[0], 3)
    rays = rays[sel].contiguous()
    m = H.module_from_field(fld, dev)
    outs = {}
    for split in (True, False):
        m.split_app = split
        o = m.render_eval(rays.to(dev), white_bg=True, want_counts=True, keep_workspace=True)
        torch.cuda.synchronize()
        outs[split] = o
    a, b = outs[True], outs[False]
    n_app = b["app_count"]
    assert int((n_app > 128).sum()) > 100 and int(((n_app > 0) & (n_app <= 128)).sum()) > 10     # both kinds of rays
    assert torch.equal(a["app_count"], n_app) and torch.equal(a["acc_map"], b["acc_map"])
    over = n_app > 128
    fa, fb = a["workspace"]["ray_feat"], b["workspace"]["ray_feat"]
    assert torch.equal(fa[over], fb[over])            # same kernel, same arithmetic
    assert float((fa - fb).abs().max()) <= 2e-6 * max(1.0, float(fb.abs().max()))
    with torch.no_grad():
        ref = orc.render_rays(fld, rays[:1024], white_bg=True)
    assert float((a["rgb_map"][:1024].cpu() - ref["rgb_map"]).abs().max()) <= TOL


@pytest.mark.parametrize("name", ["c1_dense_mask", "c2_sub", "c4_sub"])
def test_bf16_tensor_core_mlp_mode(name, dev):
    """TVM_F_MLP_BF16: the tcgen05 shade kernel — rgb within 1e-2 of the reference (north_star's bf16 MLP mode);
    depth/acc do not go through the MLP and keep the fp32 bound."""
    fld, rays, g, white, m = _case(name, dev)
    m.mlp_precision = "bf16"
    try:
        o = m.render_eval(rays.to(dev), white_bg=bool(white))
        torch.cuda.synchronize()
    finally:
        m.mlp_precision = type(m).mlp_precision
    err = np.abs(o["rgb_map"].cpu().numpy() - g["rgb_map"]).max()
    assert err <= 1e-2, err
    assert err > 0                                            # it really is the reduced-precision path
    assert np.abs(o["depth_map"].cpu().numpy() - g["depth_map"]).max() <= TOL
    assert np.abs(o["acc_map"].cpu().numpy() - g["acc_map"]).max() <= TOL


@pytest.mark.parametrize("name", ["c1_dense_mask", "c2_sub", "c4_sub"])
def test_split_operand_tensor_core_mlp_mode(name, dev):
    """TVM_F_MLP_TC3: tcgen05 shading with bf16x3 split operands — fp32-equivalent: inside the fp32 parity bound
    against the reference and within 1e-5 of the FFMA kernel."""
    fld, rays, g, white, m = _case(name, dev)
    m.mlp_precision = "fp32"
    try:
        ref = m.render_eval(rays.to(dev), white_bg=bool(white))["rgb_map"].clone()
        m.mlp_precision = "tc3"
        o = m.render_eval(rays.to(dev), white_bg=bool(white))
        torch.cuda.synchronize()
    finally:
        m.mlp_precision = type(m).mlp_precision
    assert np.abs(o["rgb_map"].cpu().numpy() - g["rgb_map"]).max() <= TOL
    assert (o["rgb_map"] - ref).abs().max().item() <= 1e-5
    assert np.abs(o["depth_map"].cpu().numpy() - g["depth_map"]).max() <= TOL


def test_forward_six_tuple_matches_oracle(dev):
    fld, rays, g, white, m = _case("c1_dense_mask", dev)
    sub = rays[3000:3700].contiguous()
    with torch.no_grad():
        rgb, depth, acc, alpha, z, dists = m(sub.to(dev), white_bg=True, is_train=False)
        o = orc.render_chunk(fld, sub, white_bg=True)
    assert alpha.shape == (700, m.nSamples) and z.shape == alpha.shape and dists.shape == alpha.shape
    assert torch.equal(z.cpu(), o["z_vals"])                     # same fp32 op order -> identical
    assert torch.equal(dists.cpu(), o["dists"])
    assert (alpha.cpu() - o["alpha"]).abs().max() <= 1e-5
    assert torch.equal(alpha.cpu() > 0, o["ray_valid"])          # sigma>0 exactly on the valid samples
    assert (rgb.cpu() - o["rgb_map"]).abs().max() <= TOL
    assert (depth.cpu() - o["depth_map"]).abs().max() <= TOL
    assert (acc.cpu() - o["acc_map"]).abs().max() <= TOL


def test_train_mode_jitter_matches_oracle(dev):
    fld, rays, g, white, m = _case("c1_dense_mask", dev)
    sub = rays[4000:4300].contiguous()
    torch.manual_seed(11)
    jit = torch.rand(300, 1)
    with torch.no_grad():
        rgb, depth, acc, alpha, z, dists = m(sub.to(dev), bg_color=torch.ones(3, device=dev), is_train=True,
                                             N_samples=443, jitter=jit.to(dev))
        o = orc.render_chunk(fld, sub, bg_color=torch.ones(3), n_samples=443, jitter=jit)
    assert torch.equal(z.cpu(), o["z_vals"])
    assert torch.equal(alpha.cpu() > 0, o["ray_valid"])
    assert (alpha.cpu() - o["alpha"]).abs().max() <= 1e-5
    assert (rgb.cpu() - o["rgb_map"]).abs().max() <= TOL


def test_octree_render_cpu_rays_and_chunking(dev):
    import iffnerf_b200 as I
    fld, rays, g, white, m = _case("c1_dense_mask", dev)
    m.max_launch_rays = 3000                                   # force several launches + the copy stream
    try:
        rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device=dev)
    finally:
        m.max_launch_rays = type(m).max_launch_rays
    assert rgb.is_cuda and rgb.shape == (10000, 3) and depth.shape == (10000,)
    assert np.abs(rgb.cpu().numpy() - g["rgb_map"]).max() <= TOL
    assert np.abs(depth.cpu().numpy() - g["depth_map"]).max() <= TOL
    # out_host: pinned result buffers filled slice by slice behind the kernels of the next slice (and on the
    # single-launch path), identical to the returned device tensors
    rgb_h, depth_h = torch.full((10000, 3), -1.0).pin_memory(), torch.full((10000,), -1.0).pin_memory()
    for cap in (3000, type(m).max_launch_rays):
        m.max_launch_rays = cap
        try:
            rgb2, _, depth2, _, _ = I.OctreeRender_trilinear_fast(rays, m, chunk=4096, white_bg=True, device=dev,
                                                              out_host=(rgb_h.fill_(-1.0), depth_h.fill_(-1.0)))
            torch.cuda.current_stream(dev).synchronize()
        finally:
            m.max_launch_rays = type(m).max_launch_rays
        assert torch.equal(rgb_h, rgb2.cpu()) and torch.equal(depth_h, depth2.cpu()) and torch.equal(rgb2, rgb)
    with pytest.raises(ValueError):
        I.OctreeRender_trilinear_fast(rays, m, white_bg=True, device=dev, out_host=(rgb_h[:5], depth_h))


def test_edge_cases(dev):
    fld, rays, g, white, m = _case("c1_dense_mask", dev)
    # empty batch
    o = m.render_eval(rays[:0].to(dev), white_bg=True)
    assert o["rgb_map"].shape == (0, 3)
    # ragged batch sizes (not multiples of the 32-ray CTA tile / 64-ray shade tile)
    for n in (1, 31, 33, 65, 127):
        o = m.render_eval(rays[5000:5000 + n].to(dev), white_bg=True)
        assert np.abs(o["rgb_map"].cpu().numpy() - g["rgb_map"][5000:5000 + n]).max() <= TOL
    # rays that miss the box, axis-parallel rays (zero direction components), origin inside the box
    special = torch.tensor([[0.0, 0.0, 4.0, 0.0, 1.0, 0.0],      # parallel to a face, misses
                            [0.0, 0.0, 4.0, 0.0, 0.0, -1.0],     # straight down the z axis
                            [0.2, -0.1, 0.3, 0.6, 0.0, -0.8],    # starts inside
                            [9.0, 9.0, 9.0, 1.0, 0.0, 0.0]])     # far away, pointing off
    fld2 = fx.make_field([128] * 3, density_shift=0.0, near_far=(0.05, 6.0), holes_seed=1)
    m2 = H.module_from_field(fld2, dev)
    o = m2.render_eval(special.to(dev), white_bg=True, early_term=False)
    ref = orc.render_chunk(fld2, special, white_bg=True)
    bits, counts = m2.sample_mask(special.to(dev))
    assert np.array_equal(H.unpack_bits(bits.cpu().numpy(), m2.nSamples), ref["ray_valid"].numpy())
    assert (o["rgb_map"].cpu() - ref["rgb_map"]).abs().max() <= TOL
    assert (o["depth_map"].cpu() - ref["depth_map"]).abs().max() <= TOL
    # N_samples override
    o = m.render_eval(rays[5000:5100].to(dev), white_bg=True, N_samples=200)
    ref = orc.render_chunk(fld, rays[5000:5100], white_bg=True, n_samples=200)
    assert (o["rgb_map"].cpu() - ref["rgb_map"]).abs().max() <= TOL


def test_pack_roundtrip_and_layout(dev, built_lib):
    """tvm_pack_factors writes the documented channel-last layout; tvm_unpack_factor_grads inverts it."""
    import ctypes as C
    from iffnerf_b200 import _lib
    fld = fx.make_field([37, 41, 29], occupancy=None)
    m = H.module_from_field(fld, dev)
    packed = m.packed_factors()
    d, host = H.host_pack_factors(m)
    assert np.array_equal(packed.cpu().numpy(), host)
    planes, lines = m._factor_params()
    outs_p = [torch.zeros_like(p) for p in planes]
    outs_l = [torch.zeros_like(p) for p in lines]
    for acc_flag, scale in ((0, 1.0), (1, 2.0)):
        _lib.check(built_lib.tvm_unpack_factor_grads(C.byref(d), _lib.ptr(packed), _lib.ptr_array(outs_p),
                                                     _lib.ptr_array(outs_l), acc_flag, None), "unpack")
        torch.cuda.synchronize()
        for a, b in zip(outs_p + outs_l, planes + lines):
            assert torch.equal(a, scale * b.detach())


def test_full_size_properties(dev):
    """BASELINE config 2 at full size (800x800, 300^3): size-independent properties.
    (a) the ray-subset golden rows are reproduced inside the full render, (b) rendering is permutation-
    equivariant (rays are independent units), (c) outputs are finite and in range, (d) acc in [0,1]."""
    fld, rays, g, white, m = _case("c2_sub", dev)
    full = fx.config2_rays()
    o = m.render_eval(full.to(dev), white_bg=True)
    rgb = o["rgb_map"]
    assert torch.isfinite(rgb).all() and rgb.min() >= 0 and rgb.max() <= 1
    assert o["acc_map"].min() >= 0 and o["acc_map"].max() <= 1 + 1e-5
    idx = torch.from_numpy(g["ray_index"])
    assert np.abs(rgb[idx.to(dev)].cpu().numpy() - g["rgb_map"]).max() <= TOL
    perm = torch.randperm(full.shape[0], generator=torch.Generator().manual_seed(0))
    o2 = m.render_eval(full[perm].to(dev), white_bg=True)
    assert torch.equal(o2["rgb_map"], rgb[perm.to(dev)])
    assert torch.equal(o2["depth_map"], o["depth_map"][perm.to(dev)])


@pytest.mark.parametrize("name,build,width", [("c2_full", lambda: fx.config2(), 800),
                                              ("c4_full", lambda: fx.config4(), 1920)])
def test_full_size_render_matches_reference_golden(name, build, width, dev):
    """BASELINE configs 2 (800x800 = 640 000 rays) and 4 (1920x1080 = 2 073 600 rays) rendered whole through
    OctreeRender_trilinear_fast against the UNMODIFIED reference's full-size render (oracle/make_golden.py
    full_image_case): every ray's valid-sample count bit-exact, a ~65 k-ray strided subset of rgb / depth within
    1e-4, and the float64 checksum of every image row within the same per-pixel tolerance."""
    import iffnerf_b200 as I
    _cache.clear()
    fld, rays = build()
    g = H.golden(name)
    H.check_params(fld, g)
    m = H.module_from_field(fld, dev)
    assert m.nSamples == int(g["n_samples"])
    rgb, _, depth, _, _ = I.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True,
                                                        ndc_ray=False, device=dev)
    counts = torch.cat([m.sample_mask(rays[a:a + (1 << 19)].to(dev), want_bits=False)[1]
                        for a in range(0, rays.shape[0], 1 << 19)])
    torch.cuda.synchronize()
    assert np.array_equal(counts.cpu().numpy().astype(np.int16), g["valid_count"])
    idx = torch.from_numpy(g["ray_index"]).to(dev)
    assert np.abs(rgb[idx].cpu().numpy() - g["rgb_sub"]).max() <= TOL
    assert np.abs(depth[idx].cpu().numpy() - g["depth_sub"]).max() <= TOL
    rows = rays.shape[0] // width
    row_rgb = rgb.double().view(rows, width, 3).sum((1, 2)).cpu().numpy()
    row_dep = depth.double().view(rows, width).sum(1).cpu().numpy()
    # row checksums: the MEAN error over a row stays well inside the per-pixel tolerance (depth carries the one-sided
    # bias of early termination: the dropped tail is <= early_term_eps * z per lit ray)
    assert np.abs(row_rgb - g["row_rgb_sum"]).max() <= TOL * 3 * width * 0.1
    assert np.abs(row_dep - g["row_depth_sum"]).max() <= TOL * width * 0.4
    del m
    torch.cuda.empty_cache()


@pytest.mark.parametrize("cfg", ["llff_like_16_4_4", "small_8_24_relu_nope"])
def test_generic_channel_counts_and_modes(cfg, dev):
    """The run-time-channel-count kernels (not the 16/48 specialisation): per-plane component counts of
    configs/flower.txt ([16,4,4] / [48,12,12]), a G=2 configuration, relu density activation, fea_pe=view_pe=0 —
    forward parity against the oracle and gradient parity against the oracle's autograd."""
    torch.manual_seed(77)
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3])
    if cfg == "llff_like_16_4_4":
        kw = dict(n_sigma=(16, 4, 4), n_app=(48, 12, 12), app_dim=27, feature_c=128, view_pe=2, fea_pe=2)
        scal = dict(density_shift=0.0, fea2dense="softplus")
    else:
        kw = dict(n_sigma=(8, 8, 8), n_app=(24, 24, 24), app_dim=27, feature_c=128, view_pe=0, fea_pe=0)
        scal = dict(density_shift=-10.0, fea2dense="relu")
    fld = orc.init_field(aabb, [40, 36, 44], near_far=[2.0, 6.0], step_ratio=0.5, distance_scale=25.0,
                         weight_thres=1e-4, scale=0.3, **scal, **kw)
    fld.occupancy = fx.sphere_occupancy(aabb, (30, 34, 38), radius=1.1, holes_seed=2)
    m = H.module_from_field(fld, dev)
    rays = fx.config1(0.0, None, 7)[1][2000:2600].contiguous()
    o = m.render_eval(rays.to(dev), white_bg=True, early_term=False, want_counts=True)
    bits, _ = m.sample_mask(rays.to(dev))
    with torch.no_grad():
        ref = orc.render_chunk(fld, rays, white_bg=True)
    assert np.array_equal(H.unpack_bits(bits.cpu().numpy(), m.nSamples), ref["ray_valid"].numpy())
    assert ref["app_mask"].sum() > 100                                       # the appearance path is exercised
    assert (o["rgb_map"].cpu() - ref["rgb_map"]).abs().max() <= TOL
    assert (o["depth_map"].cpu() - ref["depth_map"]).abs().max() <= TOL
    # gradients (march + shade backward) vs oracle autograd
    params_ref = fld.params()
    for p in params_ref:
        p.requires_grad_(True)
    torch.manual_seed(5)
    target = torch.rand(rays.shape[0], 3)
    out = orc.render_chunk(fld, rays, bg_color=torch.ones(3))
    torch.mean((out["rgb_map"] - target) ** 2).backward()
    m.zero_grad()
    rgb = m(rays.to(dev), bg_color=torch.ones(3, device=dev), is_train=False)[0]
    torch.mean((rgb - target.to(dev)) ** 2).backward()
    mine = ([*m.density_plane, *m.density_line, *m.app_plane, *m.app_line, m.basis_mat.weight]
            + [m.renderModule.mlp[i].weight for i in (0, 2, 4)] + [m.renderModule.mlp[i].bias for i in (0, 2, 4)])
    for a, b in zip(mine, params_ref):
        scale = b.grad.abs().max().item()
        if scale == 0:
            assert a.grad is None or a.grad.abs().max().item() == 0
            continue
        assert (a.grad.cpu() - b.grad).abs().max().item() <= 3e-3 * scale
    for p in params_ref:
        p.requires_grad_(False)
        p.grad = None


@pytest.mark.parametrize("mlp", ["auto", "fp32"])
def test_default_wide_head_matches_reference_golden(mlp, dev):
    """fea_pe = view_pe = 6 (the reference's default, opt.py:131-133; in_mlpC = 390): 'auto' must pick the tensor-core
    kernel (K-chunked bf16x3 variant, csrc/shade_tc3.cu::shade_tc3k_kernel) and match the reference's render to 1e-4;
    the fp32 SIMT kernel is checked against the same golden."""
    from iffnerf_b200 import synthetic as syn
    g = H.golden("c1_pe66")
    m = syn.build_model([128] * 3, dev, view_pe=6, fea_pe=6)
    got = np.array([p.double().sum().item() for p in m.state_dict().values()])
    np.testing.assert_allclose(got, g["param_checksum"], rtol=1e-10, atol=1e-10)
    assert m.renderModule.in_mlpC == int(g["in_mlpC"]) == 390
    _, rays = fx.config1(0.0, None, 7)
    rays = rays[torch.from_numpy(g["ray_index"])].contiguous().to(dev)
    m.mlp_precision = mlp
    assert m._shade_mode() == ("tc3" if mlp == "auto" else "fp32")
    o = m.render_eval(rays, white_bg=True)
    torch.cuda.synchronize()
    assert np.abs(o["rgb_map"].cpu().numpy() - g["rgb_map"]).max() <= TOL
    assert np.abs(o["depth_map"].cpu().numpy() - g["depth_map"]).max() <= TOL
