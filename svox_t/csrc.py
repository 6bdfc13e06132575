"""``svox_t.csrc`` -> the ctypes mirror of the reference's pybind11 module (svox_t_b200/csrc/__init__.py)."""
from svox_t_b200.csrc import *       # noqa: F401,F403
from svox_t_b200.csrc import (RaysSpec, TreeSpec, CameraSpec, RenderOptions, query_vertical, construct_tree,  # noqa: F401
                              volume_render, volume_render_backward, volume_render_image,
                              volume_render_image_backward, render_depth, warp_vertices, warp_vertices_backward, p2v,
                              p2v_backward, opacity_render, opacity_render_backward, motion_render)
