"""B200-native TensoRF-VM ray renderer behind the IFFNeRF API.

    from iffnerf_b200 import TensorVMSplit, AlphaGridMask, OctreeRender_trilinear_fast
"""
from .tensorf import AlphaGridMask, MLPRender_Fea, TensorVMSplit, positional_encoding  # noqa: F401
from .renderer import OctreeRender_trilinear_fast  # noqa: F401
from .raygen import pixel_rays, pixel_rays_lie  # noqa: F401
from . import sharding  # noqa: F401
from . import graphs  # noqa: F401

__all__ = ["TensorVMSplit", "AlphaGridMask", "MLPRender_Fea", "OctreeRender_trilinear_fast", "positional_encoding",
           "pixel_rays", "pixel_rays_lie"]
