"""svox_t_b200 -- B200-native (sm_100a) implementation of the octree volume-rendering hot path of
HaiminLuo/svox_t, behind svox_t's own Python API (reference: svox_t/__init__.py:30-35).

Importing the package never touches a CPU fallback: the kernels live in ``svox_t_b200/csrc/libsvoxb.so``
(C ABI in include/svoxb.h) and every operator raises if that library or a CUDA device is missing.
"""
from .version import __version__
from .svox import N3Tree, get_transformation_matrix, warp_vertices, blend_transformation_matrix
from .renderer import VolumeRenderer, NDCConfig, Rays
from .helpers import N3TreeView, LocalIndex, DataFormat
from .p2v import voxelize

__all__ = ["N3Tree", "VolumeRenderer", "NDCConfig", "Rays", "N3TreeView", "LocalIndex", "DataFormat",
           "get_transformation_matrix", "warp_vertices", "blend_transformation_matrix", "voxelize", "__version__"]
