"""TEST INFRASTRUCTURE ONLY — pins the oracle against the UNMODIFIED reference and writes tests/golden/*.npz.

Run in the authoring container (needs /root/reference):   python -m oracle.make_golden
For every case it (1) builds the seeded fixture, (2) constructs the reference's own
TensorVMSplit with the same seed and checks the parameters are identical, (3) runs the
reference's renderer and the oracle restatement on the same inputs and asserts they are
BIT-IDENTICAL, (4) stores the *reference's* outputs as the golden vectors.
"""
import contextlib
import io
import math
import os
import sys
import time

import numpy as np
import torch

from . import fixtures as fx
from . import tensorf_oracle as orc
from .ref_import import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference_model(R, fld, seed=fx.SEED):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.TensorVMSplit(fld.aabb.clone(), list(fld.grid), "cpu", density_n_comp=[16] * 3,
                            appearance_n_comp=[48] * 3, app_dim=27, near_far=list(fld.near_far),
                            shadingMode="MLP_Fea", alphaMask_thres=1e-4, density_shift=fld.density_shift,
                            distance_scale=25, pos_pe=6, view_pe=2, fea_pe=2, featureC=128, step_ratio=0.5,
                            fea2denseAct="softplus")
    sd = m.state_dict()
    mine = {}
    for k in range(3):
        mine[f"density_plane.{k}"] = fld.density_plane[k]
        mine[f"density_line.{k}"] = fld.density_line[k]
        mine[f"app_plane.{k}"] = fld.app_plane[k]
        mine[f"app_line.{k}"] = fld.app_line[k]
    mine["basis_mat.weight"] = fld.basis
    for i, li in enumerate((0, 2, 4)):
        mine[f"renderModule.mlp.{li}.weight"] = fld.mlp_w[i]
        mine[f"renderModule.mlp.{li}.bias"] = fld.mlp_b[i]
    assert set(sd.keys()) == set(mine.keys()), (sd.keys(), mine.keys())
    for k, v in sd.items():
        assert torch.equal(v, mine[k]), f"param {k} differs between reference ctor and oracle.init_field"
    if fld.occupancy is not None:
        m.alphaMask = R.AlphaGridMask("cpu", fld.occupancy.aabb.clone(), fld.occupancy.volume.clone())
    geo = orc.step_geometry(fld.aabb, fld.grid, fld.step_ratio)
    assert geo["nSamples"] == m.nSamples and torch.equal(geo["stepSize"], m.stepSize)
    return m


def reference_valid_mask(m, rays, n_samples=-1):
    """ray_valid exactly as the reference computes it (tensorBase.py:820-837), via its own methods."""
    with torch.no_grad():
        xyz, z, valid = m.sample_ray(rays[:, :3], rays[:, 3:6], None, is_train=False, N_samples=n_samples)
        if m.alphaMask is not None:
            keep = m.alphaMask.sample_alpha(xyz[valid]) > 0
            bad = ~valid
            bad[valid] |= ~keep
            valid = ~bad
    return valid


def eval_case(R, name, fld, rays, white_bg=True, bg_color=None, extra=None):
    t = time.time()
    m = build_reference_model(R, fld)
    with torch.no_grad():
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=white_bg,
                                                          bg_color=bg_color, ndc_ray=False, device="cpu")
        full = m(rays[:4096], white_bg=white_bg, bg_color=bg_color, is_train=False, N_samples=-1)
        o = orc.render_rays(fld, rays, chunk=4096, white_bg=white_bg, bg_color=bg_color,
                            keys=("rgb_map", "depth_map", "acc_map", "ray_valid", "app_mask"))
    assert torch.equal(rgb, o["rgb_map"]) and torch.equal(depth, o["depth_map"]), f"{name}: oracle != reference"
    assert torch.equal(full[2], o["acc_map"][:4096])
    valid = reference_valid_mask(m, rays)
    assert torch.equal(valid, o["ray_valid"]), f"{name}: oracle mask != reference mask"
    bits = orc.pack_valid_bits(valid)
    rec = dict(rgb_map=rgb.numpy(), depth_map=depth.numpy(), acc_map=o["acc_map"].numpy(),
               valid_bits=bits.numpy().astype(np.uint32), valid_count=valid.sum(-1).numpy().astype(np.int32),
               app_count=o["app_mask"].sum(-1).numpy().astype(np.int32),
               n_samples=np.int32(m.nSamples), step_size=np.float32(m.stepSize.item()),
               param_checksum=fx.param_checksum(fld))
    if extra:
        rec.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: rays={rays.shape[0]} S={m.nSamples} valid={valid.float().mean():.3f} "
          f"app/ray={o['app_mask'].sum(-1).float().mean():.1f} rgb_mean={rgb.mean():.4f}  ({time.time()-t:.1f}s) "
          f"oracle==reference bit-exact")


def full_image_case(R, name, fld, rays, width, stride=None, n_sub=65536):
    """Full-size render by the UNMODIFIED reference (config 2: 640 000 rays, config 4: 2 073 600 rays).  Stored:
    a strided subset of rgb/depth rows, a float64 checksum of every image row (rgb and depth), and the complete
    per-ray `ray_valid` count (bit-exact quantity)."""
    t = time.time()
    m = build_reference_model(R, fld)
    n = rays.shape[0]
    counts = np.zeros(n, dtype=np.int16)
    rgbs, depths = [], []
    with torch.no_grad():
        for a in range(0, n, 4096):
            chunk = rays[a:a + 4096]
            rgb, depth, *_ = m(chunk, white_bg=True, is_train=False, N_samples=-1)
            rgbs.append(rgb)
            depths.append(depth)
            counts[a:a + 4096] = reference_valid_mask(m, chunk).sum(-1).numpy()
    rgb, depth = torch.cat(rgbs), torch.cat(depths)
    stride = stride or max(1, n // n_sub)
    idx = np.arange(0, n, stride)
    rows = n // width
    rec = dict(ray_index=idx, rgb_sub=rgb[idx].numpy(), depth_sub=depth[idx].numpy(), valid_count=counts,
               row_rgb_sum=rgb.double().view(rows, width, 3).sum((1, 2)).numpy(),
               row_depth_sum=depth.double().view(rows, width).sum(1).numpy(),
               n_samples=np.int32(m.nSamples), step_size=np.float32(m.stepSize.item()), width=np.int32(width),
               param_checksum=fx.param_checksum(fld))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: rays={n} S={m.nSamples} valid/ray={counts.mean():.1f} rgb_mean={rgb.mean():.4f} "
          f"({time.time()-t:.1f}s) reference outputs stored (subset {idx.shape[0]}, {rows} row checksums)")


def sampled_entries(t, n=256, seed=7):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, t.numel(), (n,), generator=g)
    return idx.numpy(), t.reshape(-1)[idx].numpy()


def train_case(R, name, fld, rays, n_samples):
    """Config 3: train.py:285-339 step — forward(is_train=True) + loss + backward; grads of every parameter."""
    t = time.time()
    m = build_reference_model(R, fld)
    N = rays.shape[0]
    torch.manual_seed(1234)
    jitter = torch.rand(N, 1)
    target = torch.rand(N, 3)
    torch.manual_seed(1234)                       # the reference draws rand_like([N,1]) first thing in sample_ray
    rgb, depth, acc, alpha, z, dists = m(rays, bg_color=torch.ones(3), is_train=True, N_samples=n_samples)
    loss = torch.mean((rgb - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(alpha)))
    m.zero_grad()
    loss.backward()
    # oracle on the same jitter
    for p in fld.params():
        p.requires_grad_(True)
    o = orc.render_chunk(fld, rays, bg_color=torch.ones(3), n_samples=n_samples, jitter=jitter)
    assert torch.equal(o["rgb_map"], rgb) and torch.equal(o["alpha"], alpha) and torch.equal(o["z_vals"], z), \
        f"{name}: oracle fwd != reference (jitter draw mismatch?)"
    lo = orc.train_loss(o, target)
    assert torch.equal(lo, loss)
    grads = torch.autograd.grad(lo, fld.params())
    ref_grads = [p.grad for p in ([*m.density_plane, *m.density_line, *m.app_plane, *m.app_line, m.basis_mat.weight]
                                  + [m.renderModule.mlp[i].weight for i in (0, 2, 4)]
                                  + [m.renderModule.mlp[i].bias for i in (0, 2, 4)])]
    for a, b in zip(grads, ref_grads):
        assert torch.allclose(a, b, rtol=0, atol=0) or torch.equal(a, b), f"{name}: oracle grads != reference"
    for p in fld.params():
        p.requires_grad_(False)
    rec = dict(jitter=jitter.numpy(), target=target.numpy(), rgb_map=rgb.detach().numpy(),
               acc_map=acc.detach().numpy(), depth_map=depth.numpy(), loss=np.float64(loss.item()),
               alpha_rows=alpha[:32].detach().numpy(), alpha_sum=alpha.detach().double().sum(-1).numpy(),
               n_samples=np.int32(n_samples), param_checksum=fx.param_checksum(fld))
    names = ([f"density_plane.{k}" for k in range(3)] + [f"density_line.{k}" for k in range(3)]
             + [f"app_plane.{k}" for k in range(3)] + [f"app_line.{k}" for k in range(3)] + ["basis"]
             + [f"mlp_w{i}" for i in range(3)] + [f"mlp_b{i}" for i in range(3)])
    for nme, g in zip(names, ref_grads):
        idx, val = sampled_entries(g)
        # always include the largest-magnitude entries so the comparison is not all zeros
        top = torch.topk(g.abs().reshape(-1), min(256, g.numel())).indices
        rec[f"g_idx/{nme}"] = np.concatenate([idx, top.numpy()])
        rec[f"g_val/{nme}"] = np.concatenate([val, g.reshape(-1)[top].numpy()])
        rec[f"g_sum/{nme}"] = np.float64(g.double().sum().item())
        rec[f"g_l2/{nme}"] = np.float64(g.double().norm().item())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    # complete gradient vectors of the small parameters (lines, basis, MLP): every entry is checked, small ones included
    full = {"loss": np.float64(loss.item())}
    for nme, g in zip(names, ref_grads):
        if "plane" not in nme:
            full[f"g_full/{nme}"] = g.detach().numpy().astype(np.float32)
    np.savez_compressed(os.path.join(OUT, name + "_full.npz"), **full)
    print(f"[golden] {name}: rays={N} S={n_samples} loss={loss.item():.6f} ({time.time()-t:.1f}s) "
          f"oracle==reference bit-exact fwd+bwd")


def pose_case(R, name, fld, rays):
    """Config 5: frozen factors, gradients w.r.t. the rays (inerf/estimate_pose_inerf.py:164-178)."""
    t = time.time()
    m = build_reference_model(R, fld)
    for p in m.parameters():
        p.requires_grad_(False)
    torch.manual_seed(4321)
    bg = torch.rand(3)
    target = torch.rand(rays.shape[0], 3)
    r = rays.clone().requires_grad_(True)
    rgb, _, acc, _, _, _ = m(r, bg_color=bg, is_train=False)
    loss = torch.mean((rgb - target) ** 2)
    loss.backward()
    r2 = rays.clone().requires_grad_(True)
    o = orc.render_chunk(fld, r2, bg_color=bg)
    lo = torch.mean((o["rgb_map"] - target) ** 2)
    lo.backward()
    assert torch.equal(o["rgb_map"], rgb) and torch.equal(r.grad, r2.grad), f"{name}: oracle != reference"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), bg=bg.numpy(), target=target.numpy(),
                        rgb_map=rgb.detach().numpy(), acc_map=acc.detach().numpy(), d_rays=r.grad.numpy(),
                        loss=np.float64(loss.item()), param_checksum=fx.param_checksum(fld))
    print(f"[golden] {name}: rays={rays.shape[0]} |d_rays|max={r.grad.abs().max():.3e} ({time.time()-t:.1f}s) "
          f"oracle==reference bit-exact")


def point_case(R, name, fld, rays6):
    """model(rays, N_samples=20, sample_func=model.sample_point_color) and compute_alpha(points)."""
    t = time.time()
    m = build_reference_model(R, fld)
    with torch.no_grad():
        rgb, depth, acc, alpha, z, dists = m(rays6, N_samples=20, sample_func=m.sample_point_color, white_bg=True)
        o = orc.render_chunk(fld, rays6, white_bg=True, n_samples=20, point_samples=True)
        pts = rays6[:, :3] + 0.05 * rays6[:, 3:6]
        a_ref = m.compute_alpha(pts, length=m.stepSize.item())
        a_orc = orc.point_alpha(fld, pts, m.stepSize.item())
        f_ref = m.compute_densityfeature(m.normalize_coord(pts))
        af_ref = m.compute_appfeature(m.normalize_coord(pts))
        assert torch.equal(af_ref, orc.app_feature(fld, orc.normalize(fld, pts)))
    assert torch.equal(rgb, o["rgb_map"]) and torch.equal(alpha, o["alpha"]) and torch.equal(depth, o["depth_map"])
    assert z.shape == (1, 20) and torch.equal(z, o["z_vals"]) and torch.equal(a_ref, a_orc)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), rgb_map=rgb.numpy(), depth_map=depth.numpy(),
                        acc_map=acc.numpy(), alpha=alpha.numpy(), z_vals=z.numpy(), dists=dists.numpy(),
                        points=pts.numpy(), point_alpha=a_ref.numpy(), point_feature=f_ref.numpy(), point_appfeature=af_ref.numpy(),
                        param_checksum=fx.param_checksum(fld))
    print(f"[golden] {name}: rays={rays6.shape[0]} lit={(acc > 0).float().mean():.3f} ({time.time()-t:.1f}s) "
          f"oracle==reference bit-exact")


REF_KW = dict(density_n_comp=[16] * 3, appearance_n_comp=[48] * 3, app_dim=27, near_far=[2.0, 6.0],
              shadingMode="Ref", alphaMask_thres=1e-4, density_shift=0.0, distance_scale=25, pos_pe=6, view_pe=2,
              fea_pe=2, featureC=128, step_ratio=0.5, fea2denseAct="softplus")


def ref_head_case(R, name):
    """shadingMode='Ref' (configs/lego.txt:25).  No oracle restatement of this head: the goldens are
    the reference's own outputs and pin the product directly (eval render + one train step with all gradients)."""
    t = time.time()
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3])
    torch.manual_seed(fx.SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.TensorVMSplit(aabb.clone(), [128] * 3, "cpu", **REF_KW)
    occ = fx.sphere_occupancy(aabb, 200, radius=1.0, holes_seed=1)
    m.alphaMask = R.AlphaGridMask("cpu", occ.aabb.clone(), occ.volume.clone())
    _, rays = fx.config1(0.0, None, 7)
    sub, idx = fx.subsample(rays, 2048, seed=4)
    with torch.no_grad():
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(sub, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device="cpu")
    tr, tidx = fx.subsample(rays, 512, seed=5)
    torch.manual_seed(99)
    jitter = torch.rand(512, 1)
    target = torch.rand(512, 3)
    torch.manual_seed(99)
    out = m(tr, bg_color=torch.ones(3), is_train=True, N_samples=443)
    loss = torch.mean((out[0] - target) ** 2) + 0.1 * torch.mean(torch.exp(torch.abs(out[3])))
    m.zero_grad()
    loss.backward()
    rec = dict(ray_index=idx.numpy(), rgb_map=rgb.numpy(), depth_map=depth.numpy(), train_index=tidx.numpy(),
               jitter=jitter.numpy(), target=target.numpy(), train_rgb=out[0].detach().numpy(),
               loss=np.float64(loss.item()),
               param_checksum=np.array([p.double().sum().item() for p in m.state_dict().values()]))
    for nme, p in m.named_parameters():
        if p.grad is None:
            continue
        g = p.grad
        top = torch.topk(g.abs().reshape(-1), min(128, g.numel())).indices
        rec[f"g_idx/{nme}"] = top.numpy()
        rec[f"g_val/{nme}"] = g.reshape(-1)[top].numpy()
        rec[f"g_l2/{nme}"] = np.float64(g.double().norm().item())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: eval rays=2048 rgb_mean={rgb.mean():.4f}, train loss={loss.item():.6f} "
          f"({time.time()-t:.1f}s) reference outputs stored")


def wide_head_case(R, name):
    """The reference's DEFAULT positional-encoding widths (opt.py:131-133: fea_pe = view_pe = 6 -> in_mlpC = 390,
    tensorBase.py:165-183).  No oracle restatement is needed: the golden is the reference's own render and pins the
    product directly (its K-chunked tensor-core shading kernel and the fp32 SIMT one)."""
    t = time.time()
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3])
    torch.manual_seed(fx.SEED)
    kw = dict(REF_KW)
    kw.update(shadingMode="MLP_Fea", view_pe=6, fea_pe=6)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.TensorVMSplit(aabb.clone(), [128] * 3, "cpu", **kw)
    occ = fx.sphere_occupancy(aabb, 200, radius=1.0)
    m.alphaMask = R.AlphaGridMask("cpu", occ.aabb.clone(), occ.volume.clone())
    _, rays = fx.config1(0.0, None, 7)
    sub, idx = fx.subsample(rays, 4096, seed=6)
    with torch.no_grad():
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(sub, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device="cpu")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ray_index=idx.numpy(), rgb_map=rgb.numpy(),
                        depth_map=depth.numpy(), in_mlpC=np.int32(m.renderModule.in_mlpC),
                        param_checksum=np.array([p.double().sum().item() for p in m.state_dict().values()]))
    print(f"[golden] {name}: in_mlpC={m.renderModule.in_mlpC} rays=4096 rgb_mean={rgb.mean():.4f} "
          f"({time.time()-t:.1f}s) reference outputs stored")


def raygen_case(R, name):
    """rays of 1024 random pixels through a perturbed orbit pose, exactly as the iNeRF loop builds them
    (inerf/estimate_pose_inerf.py:96-99,149-164), and the gradient of a fixed linear functional w.r.t. the pose."""
    import torch.nn.functional as F
    t = time.time()
    H = W = 800
    focal = 400.0 / math.tan(0.5 * 0.6911112)
    K = torch.tensor([[[focal, 0.0, W / 2], [0.0, focal, H / 2], [0.0, 0.0, 1.0]]])
    g = torch.Generator().manual_seed(55176280)
    c2w = torch.cat([fx.orbit_pose(), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
    c2w = (c2w + 0.01 * torch.randn(4, 4, generator=g) * torch.tensor([[1.0], [1.0], [1.0], [0.0]])).contiguous()
    pixels = torch.stack([torch.randint(0, W, (1024,), generator=g), torch.randint(0, H, (1024,), generator=g)], -1)
    upstream = torch.randn(1024, 7, generator=g)
    pose = c2w.clone().requires_grad_(True)
    ori, dx, dy = R.get_ray_directions_Ks(H, W, K, use_pixel_centers=True)
    directions = ori / torch.linalg.norm(ori, dim=-1, keepdim=True)
    ro, rd, radii = R.get_rays(directions, pose, directions=ori, dx=dx, dy=dy, keepdim=True)
    bx, by = pixels[:, 0], pixels[:, 1]
    rays = torch.cat((ro[0, by, bx], F.normalize(rd[0, by, bx], p=2, dim=-1), radii[0, by, bx]), dim=-1)
    (rays * upstream).sum().backward()
    pose2 = c2w.clone().requires_grad_(True)
    mine = orc.pixel_rays(K, pose2, pixels, H, W)
    (mine * upstream).sum().backward()
    assert torch.equal(mine, rays) and torch.equal(pose.grad, pose2.grad), f"{name}: oracle != reference"
    with torch.no_grad():       # loader semantics (dataLoader/blender.py:105-114): no second normalisation
        full = torch.cat(R.get_rays(directions, c2w, directions=ori, dx=dx, dy=dy, keepdim=True), -1)[0]
        loader = full[by, bx]
        assert torch.equal(loader, orc.pixel_rays(K, c2w, pixels, H, W, renormalize=False))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), K=K.numpy(), c2w=c2w.numpy(), pixels=pixels.numpy(),
                        upstream=upstream.numpy(), rays=rays.detach().numpy(), loader_rays=loader.numpy(),
                        d_c2w=pose.grad.numpy(), hw=np.array([H, W]))
    print(f"[golden] {name}: 1024 pixels, |d_c2w|max={pose.grad.abs().max():.3e} ({time.time()-t:.1f}s) "
          f"oracle==reference bit-exact fwd+bwd")


GRIDOPS_GRID = [32, 28, 36]
GRIDOPS_LATTICE = (48, 40, 56)
GRIDOPS_UPSAMPLE = [44, 38, 50]


def gridops_fixture():
    """Small non-cubic truck-shaped field + 960 rays (some miss the box) for the grid-maintenance / checkpoint cases."""
    aabb = torch.tensor(fx.TRUCK_AABB)
    fld = fx.make_field(GRIDOPS_GRID, aabb=aabb, near_far=(0.01, 6.0), occ_res=(30, 34, 26))
    c2w = fx.look_at_c2w((2.2, 1.6, 0.9), target=(0.0, 0.0, 0.25))
    rays = fx.pinhole_rays(24, 40, 0.5 * 40, c2w, cols=7)
    return fld, rays


def _state_checksums(m):
    return {k: np.array([v.double().sum().item(), v.double().abs().sum().item()]) for k, v in m.state_dict().items()}


def gridops_case(R, name, ckpt_name):
    """SURVEY 8f-3 / 8f-4 on the UNMODIFIED reference: updateAlphaMask on a (48,40,56) lattice (threshold placed in the
    widest gap of the pooled alpha values near their median, so the mask is robust to fp32 summation order),
    filtering_rays in both modes, shrink, upsample_volume_grid, and `save()` of the resulting model; each step's
    observable results are stored (mask bits, boxes, kept-ray masks, grid sizes, factor checksums, renders)."""
    import torch.nn.functional as F
    t = time.time()
    fld, rays = gridops_fixture()
    m = build_reference_model(R, fld)
    rec = dict(param_checksum=fx.param_checksum(fld))
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        alpha, _ = m.getDenseAlpha(GRIDOPS_LATTICE)
        pooled = F.max_pool3d(alpha.clamp(0, 1).transpose(0, 2).contiguous()[None, None], kernel_size=3, padding=1,
                              stride=1).reshape(-1)
        vals = torch.sort(pooled[pooled > 0]).values
        mid = vals.numel() // 2
        window = vals[mid - 400:mid + 400]
        gaps = window[1:] - window[:-1]
        j = int(torch.argmax(gaps))
        thres = float((window[j].double() + window[j + 1].double()) / 2)
        m.alphaMask_thres = thres
        new_aabb = m.updateAlphaMask(GRIDOPS_LATTICE)
        vol = m.alphaMask.alpha_volume.reshape(GRIDOPS_LATTICE[::-1])
        rec.update(thres=np.float64(thres), thres_margin=np.float64(float(gaps[j]) / 2),
                   mask_bits=np.packbits(vol.bool().numpy().reshape(-1)), mask_shape=np.array(vol.shape),
                   mask_aabb=new_aabb.numpy(), mask_fraction=np.float64(vol.mean().item()))
        index = torch.arange(rays.shape[0])[:, None].float()
        for mode, bbox_only in (("box", True), ("occ", False)):
            _, kept = m.filtering_rays(rays, index, N_samples=64, bbox_only=bbox_only)
            keep = torch.zeros(rays.shape[0], dtype=torch.bool)
            keep[kept.reshape(-1).long()] = True
            rec[f"filter_{mode}"] = keep.numpy()
        m.shrink(new_aabb)
        rec.update(shrink_grid=np.array(m.gridSize.tolist()), shrink_aabb=m.aabb.numpy().copy(),
                   shrink_step=np.float32(m.stepSize.item()), shrink_nsamples=np.int32(m.nSamples))
        for k, v in _state_checksums(m).items():
            rec[f"shrink_sum/{k}"] = v
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device="cpu")
        rec.update(shrink_rgb=rgb.numpy(), shrink_depth=depth.numpy(),
                   shrink_valid_count=reference_valid_mask(m, rays).sum(-1).numpy().astype(np.int32))
        m.upsample_volume_grid(GRIDOPS_UPSAMPLE)
        rec.update(up_grid=np.array(m.gridSize.tolist()), up_step=np.float32(m.stepSize.item()),
                   up_nsamples=np.int32(m.nSamples))
        for k, v in _state_checksums(m).items():
            rec[f"up_sum/{k}"] = v
        for k in range(3):      # a few exact entries of every resized factor
            for nme, p in ((f"density_plane.{k}", m.density_plane[k]), (f"app_line.{k}", m.app_line[k])):
                idx, val = sampled_entries(p.data, n=128, seed=3 + k)
                rec[f"up_idx/{nme}"] = idx
                rec[f"up_val/{nme}"] = val
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device="cpu")
        rec.update(up_rgb=rgb.numpy(), up_depth=depth.numpy(),
                   up_valid_count=reference_valid_mask(m, rays).sum(-1).numpy().astype(np.int32))
        m.save(os.path.join(OUT, ckpt_name))                  # the reference's own `.th` (tensorBase.py:424-442)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
    print(f"[golden] {name}: lattice {GRIDOPS_LATTICE} occupied {rec['mask_fraction']:.3f} (threshold margin "
          f"{rec['thres_margin']:.2e}), kept rays box/occ {rec['filter_box'].sum()}/{rec['filter_occ'].sum()} of "
          f"{rays.shape[0]}, shrink -> {rec['shrink_grid'].tolist()}, upsample -> {rec['up_grid'].tolist()}, "
          f"checkpoint {ckpt_name} ({os.path.getsize(os.path.join(OUT, ckpt_name)) >> 10} KiB)  ({time.time()-t:.1f}s)")


def check_product_checkpoint(R, path, rays):
    """Reverse direction of the `.th` contract: a checkpoint written by the PRODUCT's save() is loaded by the
    unmodified reference (the loader of train.py:40-45 / pose_estimation/model_utils.py:6-12) and must reproduce the
    state and the render of the reference model it was copied from."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    kwargs = dict(ckpt["kwargs"])
    kwargs.update(device="cpu")
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.TensorVMSplit(**kwargs)
        m.load(ckpt)
    with torch.no_grad():
        rgb, _, depth, _, _ = R.OctreeRender_trilinear_fast(rays, m, chunk=4096, N_samples=-1, white_bg=True,
                                                          ndc_ray=False, device="cpu")
    return m, rgb, depth


def point_rays(fld, n, seed=11):
    """Points near the occupied shell with isocell-like random directions (6-col rays)."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(n, 3, generator=g)
    p = p / p.norm(dim=-1, keepdim=True) * (0.55 + 0.5 * torch.rand(n, 1, generator=g))
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return torch.cat([p, d], -1).contiguous()


def main():
    os.makedirs(OUT, exist_ok=True)
    R = import_reference()
    torch.set_num_threads(os.cpu_count())
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    if want("c1_dense_mask"):
        fld, rays = fx.config1(density_shift=0.0, occupancy="sphere", cols=6)
        eval_case(R, "c1_dense_mask", fld, rays)
    if want("c1_refdefault_nomask"):
        fld, rays = fx.config1(density_shift=-10.0, occupancy=None, cols=6)
        eval_case(R, "c1_refdefault_nomask", fld, rays)
    if want("c1_dense_7col_blackbg"):
        fld, rays = fx.config1(density_shift=0.0, occupancy="sphere", cols=7)
        rays, idx = fx.subsample(rays, 2048, seed=3)
        eval_case(R, "c1_dense_7col_blackbg", fld, rays, white_bg=None, extra=dict(ray_index=idx.numpy()))
    if any(want(n) for n in ("c2_sub", "c3_train", "c5_pose")):
        fld, rays = fx.config2()
        if want("c2_sub"):
            sub, idx = fx.subsample(rays, 2048, seed=0)
            eval_case(R, "c2_sub", fld, sub, extra=dict(ray_index=idx.numpy()))
        if want("c3_train"):
            sub, idx = fx.subsample(rays, 1024, seed=1)
            train_case(R, "c3_train", fld, sub, n_samples=1039)
        if want("c5_pose"):
            sub, idx = fx.subsample(rays, 512, seed=2)
            pose_case(R, "c5_pose", fld, sub)
    if want("c1_point20"):
        fld, _ = fx.config1(density_shift=0.0, occupancy="sphere", cols=6)
        point_case(R, "c1_point20", fld, point_rays(fld, 4096))
    if want("c1_ref_head"):
        ref_head_case(R, "c1_ref_head")
    if want("c5_raygen"):
        raygen_case(R, "c5_raygen")
    if "c2_full" in only:        # minutes of CPU: only on request
        fld, rays = fx.config2()
        full_image_case(R, "c2_full", fld, rays, 800)
    if "c4_full" in only:
        fld, rays = fx.config4()
        full_image_case(R, "c4_full", fld, rays, 1920)
    if want("c1_pe66"):
        wide_head_case(R, "c1_pe66")
    if want("c_gridops"):
        gridops_case(R, "c_gridops", "c_ckpt_reference.th")
    if want("c4_sub"):
        fld, rays = fx.config4()
        sub, idx = fx.subsample(rays, 2048, seed=0)
        eval_case(R, "c4_sub", fld, sub, extra=dict(ray_index=idx.numpy(), grid=np.array(fld.grid)))


if __name__ == "__main__":
    main()
