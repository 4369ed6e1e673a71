"""System test: the reference's training schedule (train.py:262-415) end to end on the product —
Adam over get_optparam_groups, the L1 / orthogonality regularisers and the alpha loss of train.py:293-329,
updateAlphaMask + shrink (:384-395), filtering_rays (:397-400), upsample_volume_grid + optimiser rebuild (:402-415)
and an evaluation render — on a small synthetic scene whose ground truth is rendered by a frozen "teacher" field."""
import contextlib
import io
import math

import numpy as np
import pytest
import torch

from oracle import fixtures as fx

pytestmark = pytest.mark.gpu


def _model(dev, reso, seed, density_shift):
    import iffnerf_b200 as I
    aabb = torch.tensor([[-1.5] * 3, [1.5] * 3], device=dev)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return I.TensorVMSplit(aabb, [reso] * 3, dev, density_n_comp=[16] * 3, appearance_n_comp=[48] * 3, app_dim=27,
                               near_far=[2.0, 6.0], shadingMode="MLP_Fea", alphaMask_thres=1e-4,
                               density_shift=density_shift, distance_scale=25, pos_pe=6, view_pe=2, fea_pe=2,
                               featureC=128, step_ratio=0.5, fea2denseAct="softplus")


def _psnr(mse):
    return -10.0 * math.log10(max(mse, 1e-12))


def test_reference_training_schedule_end_to_end(built_lib):
    import iffnerf_b200 as I
    dev = torch.device("cuda:0")
    # ---- ground truth: a frozen teacher field confined to a ball of radius 0.9, seen from 6 orbit views of 96x96 rays
    teacher = _model(dev, 64, 123, density_shift=-2.0)
    occ = fx.sphere_occupancy(torch.tensor([[-1.5] * 3, [1.5] * 3]), 96, radius=0.9)
    teacher.alphaMask = I.AlphaGridMask(dev, occ.aabb.to(dev), occ.volume.to(dev))
    Hh = Ww = 96
    focal = 0.5 * Ww / math.tan(0.5 * 0.6911112)
    views = [fx.pinhole_rays(Hh, Ww, focal, fx.orbit_pose(60.0 * k, 20.0 + 10.0 * (k % 2)), cols=7) for k in range(6)]
    allrays = torch.cat(views, 0).to(dev)
    with torch.no_grad():
        allrgbs, _, _, _, _ = I.OctreeRender_trilinear_fast(allrays, teacher, white_bg=True, device=dev)
    test_rays, test_rgb = allrays[: Hh * Ww], allrgbs[: Hh * Ww]
    assert 0.05 < float((allrgbs < 0.99).any(-1).float().mean()) < 0.95           # the scene is neither empty nor full

    # ---- student, trained with the reference's loop
    student = _model(dev, 48, 20211202, density_shift=-2.0)
    optimizer = torch.optim.Adam(student.get_optparam_groups(0.02, 1e-3), betas=(0.9, 0.99))
    n_iters, batch = 400, 2048
    update_alpha_mask_list, upsample_list = [150, 250], [200, 300]
    n_voxel_list = [80 ** 3, 112 ** 3]
    lr_factor = 0.1 ** (1.0 / n_iters)
    reso_cur = [48] * 3
    n_samples = student.nSamples
    gen = torch.Generator(device="cpu").manual_seed(0)
    white = torch.ones(3, device=dev)
    hist = []
    with torch.no_grad():
        rgb0, _, _, _, _ = I.OctreeRender_trilinear_fast(test_rays, student, white_bg=True, device=dev)
    psnr_start = _psnr(float(torch.mean((rgb0 - test_rgb) ** 2)))
    for it in range(n_iters):
        idx = torch.randint(0, allrays.shape[0], (batch,), generator=gen).to(dev)
        rgb_map, depth_map, acc_map, weights, z_vals, dists = student(allrays[idx], N_samples=n_samples, bg_color=white,
                                                                      ndc_ray=False, is_train=True)
        loss = torch.mean((rgb_map - allrgbs[idx]) ** 2)
        total = loss + 1e-4 * student.vector_comp_diffs() + 8e-5 * student.density_L1() \
            + 0.1 * torch.exp(weights.abs()).mean()
        optimizer.zero_grad()
        total.backward()
        optimizer.step()
        hist.append(loss.item())
        for group in optimizer.param_groups:
            group["lr"] = group["lr"] * lr_factor
        if it in update_alpha_mask_list:
            new_aabb = student.updateAlphaMask(tuple(reso_cur))
            if it == update_alpha_mask_list[0]:
                student.shrink(new_aabb)
                size = student.aabb.cpu()[1] - student.aabb.cpu()[0]
                assert torch.isfinite(size).all() and (size > 0).all() and size.max() <= 3.0 + 1e-6
            else:
                n_before = allrays.shape[0]
                allrays, allrgbs = student.filtering_rays(allrays, allrgbs)
                assert 0 < allrays.shape[0] <= n_before
        if it in upsample_list:
            from oracle import tensorf_oracle as orc
            reso_cur = orc.n_to_reso(n_voxel_list.pop(0), student.aabb.cpu())
            student.upsample_volume_grid(reso_cur)
            n_samples = student.nSamples
            optimizer = torch.optim.Adam(student.get_optparam_groups(0.02 * lr_factor ** it, 1e-3 * lr_factor ** it),
                                         betas=(0.9, 0.99))
    assert all(np.isfinite(hist))
    first, last = float(np.mean(hist[:10])), float(np.mean(hist[-10:]))
    assert last < 0.15 * first, (first, last)
    with torch.no_grad():
        rgb1, _, depth1, _, _ = I.OctreeRender_trilinear_fast(test_rays, student, white_bg=True, device=dev)
    psnr_end = _psnr(float(torch.mean((rgb1 - test_rgb) ** 2)))
    assert psnr_end > psnr_start + 8.0, (psnr_start, psnr_end)
    assert torch.isfinite(depth1).all()
    # checkpoint round trip of the trained, shrunk, upsampled model (tensorBase.py:424-458)
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "student.th")
        student.save(path)
        ckpt = torch.load(path, map_location=dev, weights_only=False)
        kw = ckpt["kwargs"]
        kw.update({"device": dev})
        with contextlib.redirect_stdout(io.StringIO()):
            again = I.TensorVMSplit(**kw)
        again.load(ckpt)
    with torch.no_grad():
        rgb2, _, _, _, _ = I.OctreeRender_trilinear_fast(test_rays, again, white_bg=True, device=dev)
    assert (rgb2 - rgb1).abs().max().item() <= 1e-5
