"""CPU-side checks: the C-ABI library builds for sm_100a, loads, and exports every symbol the header
declares; the host mirror keeps the reference's constructor/state_dict/error contract."""
import io
import os
import re
import contextlib

import numpy as np
import pytest
import torch

from oracle import fixtures as fx
from oracle import tensorf_oracle as orc
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "tvm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tvm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    from iffnerf_b200 import _lib
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/tvm_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == declared        # the ctypes binding covers the whole header
    assert built_lib.tvm_abi_version() == _lib.ABI_VERSION
    assert built_lib.tvm_error_string(-3) == b"tvm: workspace too small"


def test_library_is_sm100a_native():
    from iffnerf_b200 import build
    import subprocess, shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", build.LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_field_desc_struct_matches_header(built_lib):
    """Size of the ctypes mirror == sizeof(tvm_field_desc) as the compiler lays it out."""
    import ctypes as C, subprocess, tempfile
    from iffnerf_b200 import _lib
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "tvm_b200.h"\nint main(){printf("%zu %zu %zu %zu",'
                           'sizeof(tvm_field_desc), offsetof(tvm_field_desc, dplane_off), offsetof(tvm_field_desc, occ_cells),'
                           'offsetof(tvm_field_desc, factors));return 0;}')
        exe = os.path.join(td, "sz")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        size, o1, o2, o3 = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    assert C.sizeof(_lib.FieldDesc) == size
    assert _lib.FieldDesc.dplane_off.offset == o1
    assert _lib.FieldDesc.occ_cells.offset == o2
    assert _lib.FieldDesc.factors.offset == o3


def test_scatter_out_and_ref_head_structs_match_header(built_lib):
    """tvm_scatter_out / tvm_ref_head: ctypes mirrors laid out exactly as the compiler lays out the header's structs."""
    import ctypes as C, subprocess, tempfile
    from iffnerf_b200 import _lib
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "sz.c")
        open(c, "w").write('#include <stdio.h>\n#include <stddef.h>\n#include "tvm_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %d",'
                           'sizeof(tvm_scatter_out), offsetof(tvm_scatter_out, n_dst), offsetof(tvm_scatter_out, rgb),'
                           'offsetof(tvm_scatter_out, depth), sizeof(tvm_ref_head), TVM_MAX_PEERS);return 0;}')
        exe = os.path.join(td, "sz")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        size, o1, o2, o3, ref_size, peers = map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split())
    assert C.sizeof(_lib.ScatterOut) == size and _lib.MAX_PEERS == peers
    assert _lib.ScatterOut.n_dst.offset == o1 and _lib.ScatterOut.rgb.offset == o2 and _lib.ScatterOut.depth.offset == o3
    assert C.sizeof(_lib.RefHead) == ref_size


def _module(grid=(24, 20, 28), **kw):
    import iffnerf_b200 as I
    args = dict(density_n_comp=[16] * 3, appearance_n_comp=[48] * 3, app_dim=27, shadingMode="MLP_Fea",
                view_pe=2, fea_pe=2, step_ratio=0.5, density_shift=0.0)
    args.update(kw)
    with contextlib.redirect_stdout(io.StringIO()):
        return I.TensorVMSplit(torch.tensor([[-1.5, -1.2, -1.0], [1.5, 1.3, 1.1]]), list(grid), "cpu", **args)


def test_constructor_rng_order_and_state_dict_keys():
    """One seed -> the same parameters as the reference constructor order (via the pinned oracle.init_field)."""
    torch.manual_seed(fx.SEED)
    m = _module()
    torch.manual_seed(fx.SEED)
    fld = orc.init_field(m.aabb, [24, 20, 28], **fx.MODEL_KW)
    sd = m.state_dict()
    expect = ([f"density_plane.{k}" for k in range(3)] + [f"density_line.{k}" for k in range(3)]
              + [f"app_plane.{k}" for k in range(3)] + [f"app_line.{k}" for k in range(3)] + ["basis_mat.weight"]
              + [f"renderModule.mlp.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")])
    assert list(sd.keys()) == expect
    for k in range(3):
        assert torch.equal(sd[f"density_plane.{k}"], fld.density_plane[k])
        assert torch.equal(sd[f"app_line.{k}"], fld.app_line[k])
    assert torch.equal(sd["basis_mat.weight"], fld.basis)
    assert torch.equal(sd["renderModule.mlp.4.weight"], fld.mlp_w[2])
    assert sd["density_plane.1"].shape == (1, 16, 28, 24) and sd["density_line.1"].shape == (1, 16, 20, 1)


def test_step_geometry_matches_oracle():
    m = _module()
    geo = orc.step_geometry(m.aabb, [24, 20, 28], 0.5)
    assert m.nSamples == geo["nSamples"]
    assert torch.equal(m.stepSize, geo["stepSize"]) and torch.equal(m.invaabbSize, geo["invaabbSize"])
    assert m.gridSize.tolist() == [24, 20, 28]
    groups = m.get_optparam_groups(0.02, 1e-3)
    assert [g["lr"] for g in groups] == [0.02] * 4 + [1e-3] * 2


def test_checkpoint_roundtrip(tmp_path):
    import iffnerf_b200 as I
    m = _module()
    vol = (torch.rand(9, 10, 11) > 0.5).float()
    m.alphaMask = I.AlphaGridMask("cpu", m.aabb, vol)
    path = str(tmp_path / "m.th")
    m.save(path)
    ckpt = torch.load(path, weights_only=False)
    assert ckpt["model_name"] == "TensorVMSplit" and set(ckpt) >= {"kwargs", "state_dict", "alphaMask.mask"}
    kwargs = ckpt["kwargs"]
    kwargs.update({"device": "cpu"})
    with contextlib.redirect_stdout(io.StringIO()):
        m2 = I.TensorVMSplit(**kwargs)
    m2.load(ckpt)
    for a, b in zip(m.state_dict().values(), m2.state_dict().values()):
        assert torch.equal(a, b)
    assert torch.equal(m2.alphaMask.alpha_volume, m.alphaMask.alpha_volume)
    assert m2.alphaMask.gridSize.tolist() == [11, 10, 9]


def test_no_cpu_fallback_and_unsupported_modes_raise():
    import iffnerf_b200 as I
    from iffnerf_b200 import _lib
    m = _module()
    rays = torch.zeros(4, 6)
    with torch.no_grad():
        with pytest.raises(_lib.TvmError):
            m(rays)                                            # CPU rays: there is no CPU path
        with pytest.raises(NotImplementedError):
            m(rays, ndc_ray=True)
        with pytest.raises(NotImplementedError):
            m(rays, sample_func=lambda *a, **k: None)
    with pytest.raises(RuntimeError):
        I.OctreeRender_trilinear_fast(rays, m, device="cpu")
    with pytest.raises(NotImplementedError):
        _module(shadingMode="SH")
    # the ray generator and the point queries have no CPU path either
    K = torch.tensor([[[100.0, 0.0, 50.0], [0.0, 100.0, 50.0], [0.0, 0.0, 1.0]]])
    with pytest.raises(_lib.TvmError):
        I.pixel_rays(K, torch.eye(4), torch.zeros(4, 2, dtype=torch.int32))
    with pytest.raises(_lib.TvmError):
        m.compute_appfeature(torch.zeros(4, 3))
    with pytest.raises(_lib.TvmError):
        m.compute_alpha(torch.zeros(4, 3))


def test_missing_library_fails_loudly(tmp_path):
    """No CUDA library -> TvmError naming the build command; nothing falls back to another implementation."""
    import subprocess, sys
    code = ("import os, sys; os.environ['TVM_B200_LIB'] = sys.argv[1]\n"
            "from iffnerf_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.TvmError as e:\n"
            "    assert 'no CPU fallback' in str(e) and 'iffnerf_b200.build' in str(e); print('loud')\n"
            "else:\n    raise SystemExit('load() succeeded without a library')\n")
    out = subprocess.run([sys.executable, "-c", code, str(tmp_path / "absent.so")], check=True, cwd=ROOT,
                         capture_output=True, text=True).stdout
    assert "loud" in out


def test_ray_generator_argument_checks():
    import iffnerf_b200 as I
    K = torch.tensor([[[100.0, 0.0, 50.0], [0.0, 100.0, 50.0], [0.0, 0.0, 1.0]]])
    with pytest.raises(ValueError):
        I.pixel_rays(K, torch.eye(4), None)                    # neither pixels nor image_wh
    with pytest.raises(ValueError):
        I.pixel_rays(K, torch.eye(4)[None].repeat(2, 1, 1), None, image_wh=(8, 8))   # full image takes one pose
    from iffnerf_b200 import raygen
    k1 = raygen._kinv(K)
    assert raygen._kinv(K) is k1                                # cached per (tensor, version)
    ref = torch.inverse(K[0])
    assert max(abs(k1[i] - ref.reshape(-1)[i].item()) for i in range(9)) < 1e-9


def test_product_package_never_imports_oracle():
    import subprocess, sys
    code = ("import sys; import iffnerf_b200, iffnerf_b200.renderer, iffnerf_b200.tensorf; "
            "bad=[m for m in sys.modules if m.split('.')[0]=='oracle']; assert not bad, bad")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
    for root, _, files in os.walk(os.path.join(ROOT, "iffnerf_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
