"""Driver for the compiled reference extension (oracle/_ref/svox_t_ref_csrc.so, built by oracle/build_ref.sh).

TEST INFRASTRUCTURE ONLY. The reference's pybind11 module is driven directly through its own TreeSpec / RaysSpec /
RenderOptions classes (svox_t/csrc/svox.cpp:73-145) -- the reference Python package is not imported. Needs a GPU.
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "..", "oracle", "_ref", "svox_t_ref_csrc.so")
_mod = None


def available():
    return os.path.exists(REF_SO) and torch.cuda.is_available()


def module():
    global _mod
    if _mod is None:
        spec = importlib.util.spec_from_file_location("svox_t_ref_csrc", REF_SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def tree_spec(features, child, data, parent_depth, offset, scaling, n_internal, dtype=torch.float32):
    m = module()
    dev = features.device
    ts = m.TreeSpec()
    ts.features = features.to(dtype).contiguous()
    ts.data = data.contiguous()
    ts.child = child.contiguous()
    ts.parent_depth = parent_depth.contiguous()
    ts.extra_data = torch.empty((0, 0), device=dev, dtype=dtype)
    ts.offset = offset.to(dtype).contiguous()
    ts.scaling = scaling.to(dtype).contiguous()
    ts._weight_accum = torch.empty(0, device=dev, dtype=dtype)
    ts.joint_features = torch.empty((0, 0), device=dev, dtype=dtype)
    ts.skinning_weights = torch.empty((0, 0), device=dev, dtype=dtype)
    ts.joint_index = torch.empty((0, 0), device=dev, dtype=torch.int32)
    ts.transformation_matrices = torch.empty((0, 0, 0), device=dev, dtype=dtype)
    ts.n_internal = int(n_internal)
    return ts


def rays_spec(origins, dirs, dtype=torch.float32):
    m = module()
    rs = m.RaysSpec()
    rs.origins = origins.to(dtype).contiguous()
    rs.dirs = dirs.to(dtype).contiguous()
    rs.vdirs = dirs.to(dtype).contiguous()
    return rs


def options(step_size=1e-3, background_brightness=1.0, sigma_thresh=0.0, stop_thresh=0.0):
    m = module()
    o = m.RenderOptions()
    o.step_size = step_size
    o.background_brightness = background_brightness
    o.format = 0
    o.basis_dim = -1
    o.ndc_width = -1
    o.ndc_height = -1
    o.ndc_focal = 0.0
    o.min_comp = 0
    o.max_comp = -1
    o.sigma_thresh = sigma_thresh
    o.stop_thresh = stop_thresh
    return o
