"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference, from /root/reference (authoring container) or from the
byte-identical copy of its render-path files under oracle/_ref/ (built by `python -m oracle.build_ref`; travels to the
GPU box).  Used by oracle/make_golden.py to (a) pin oracle/tensorf_oracle.py against the real reference and
(b) generate tests/golden/*.npz, and by `bench.py --impl reference` to time the reference's own CPU renderer.
Nothing in the product package or the `-m gpu` tests imports this module.

The reference has import-time dependencies on packages that are absent here and
are not on the render path; they are replaced by empty stubs.
"""
import sys
import types

import os

REF_ROOT = "/root/reference"
LOCAL_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def reference_root():
    """/root/reference when present, else the travelling copy oracle/_ref, else None."""
    if os.path.isdir(REF_ROOT):
        return REF_ROOT
    if os.path.exists(os.path.join(LOCAL_REF, "renderer.py")):
        return LOCAL_REF
    return None


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference(root=None):
    """Returns a namespace with the reference's render-path symbols (root: see reference_root())."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("no reference tree: neither /root/reference nor oracle/_ref (python -m oracle.build_ref)")
    if "omegaconf" not in sys.modules:
        _stub("omegaconf", OmegaConf=object)
    if "plyfile" not in sys.modules:
        _stub("plyfile")
    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.measure = _stub("skimage.measure")
    if "imageio" not in sys.modules:
        _stub("imageio")
    if "kornia" not in sys.modules:
        k = _stub("kornia", create_meshgrid=None)
        kg = _stub("kornia.geometry")
        kl = _stub("kornia.geometry.liegroup", Se3=object)
        k.geometry = kg
        kg.liegroup = kl
        k.__path__ = []
        kg.__path__ = []
    if "dataLoader" not in sys.modules:
        dl = _stub("dataLoader")
        dl.__path__ = [root + "/dataLoader"]  # skip dataLoader/__init__.py (imports every dataset)
    if root not in sys.path:
        sys.path.insert(0, root)
    from models.tensoRF import TensorVMSplit
    from models.tensorBase import AlphaGridMask, raw2alpha, MLPRender_Fea
    from renderer import OctreeRender_trilinear_fast
    from ray_utils import get_ray_directions_Ks, get_rays
    from utils import N_to_reso, cal_n_samples
    return types.SimpleNamespace(
        TensorVMSplit=TensorVMSplit, AlphaGridMask=AlphaGridMask, raw2alpha=raw2alpha,
        MLPRender_Fea=MLPRender_Fea, OctreeRender_trilinear_fast=OctreeRender_trilinear_fast,
        get_ray_directions_Ks=get_ray_directions_Ks, get_rays=get_rays,
        N_to_reso=N_to_reso, cal_n_samples=cal_n_samples)
