"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference).
Used by oracle/make_golden.py to (a) pin oracle/tensorf_oracle.py against the real
reference and (b) generate tests/golden/*.npz.  Nothing in the product package,
bench.py or the `-m gpu` tests imports this module.

The reference has import-time dependencies on packages that are absent here and
are not on the render path; they are replaced by empty stubs.
"""
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns a namespace with the reference's render-path symbols."""
    import os
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("reference tree not present (expected only in the authoring container)")
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
        dl.__path__ = [REF_ROOT + "/dataLoader"]  # skip dataLoader/__init__.py (imports every dataset)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
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
