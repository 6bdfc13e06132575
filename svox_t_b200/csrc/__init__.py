"""``svox_t_b200.csrc`` -- host-side mirror of the reference's pybind11 module ``svox_t.csrc``
(reference: svox_t/csrc/svox.cpp:73-145) on top of the C-ABI library ``libsvoxb.so`` (include/svoxb.h).

Same names, argument meaning and error behaviour as the reference: ``TreeSpec`` / ``RaysSpec`` / ``CameraSpec`` /
``RenderOptions`` are mutable records with the reference's field names; every function takes CUDA, contiguous
torch tensors, allocates its outputs on the inputs' device (as the reference does with ``torch::zeros/empty``)
and raises ``RuntimeError`` on a failed check. Differences, all deliberate:

* kernels launch on torch's *current* stream (the reference uses the legacy default stream);
* float64 (the reference dispatches AT_DISPATCH_FLOATING_TYPES) is implemented for the hot path proper -- point query,
  ray-batch / camera render forward + backward, depth, RGBA format -- on the general kernels (``*_f64`` entry points); the
  other operators are float32 only and say so;
* there is NO CPU fallback: a missing ``libsvoxb.so`` or a non-CUDA tensor is an error, never a slow path.

PyTorch is plumbing here (device memory, streams); all compute happens in hand-written sm_100a kernels.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# SVOXB_LIBRARY: load another build of the same library (kernel experiments); the default is the in-tree one.
_SO_PATH = os.environ.get("SVOXB_LIBRARY") or os.path.join(_HERE, "libsvoxb.so")

FORMAT_RGBA, FORMAT_SH, FORMAT_SG, FORMAT_ASG = 0, 1, 2, 3


class _CTree(ctypes.Structure):
    _fields_ = [
        ("features", ctypes.c_void_p), ("M", ctypes.c_int64), ("D", ctypes.c_int32), ("N", ctypes.c_int32),
        ("child", ctypes.c_void_p), ("data", ctypes.c_void_p), ("parent_depth", ctypes.c_void_p),
        ("n_nodes", ctypes.c_int64), ("n_internal", ctypes.c_int64),
        ("offset", ctypes.c_void_p), ("scaling", ctypes.c_void_p), ("accel", ctypes.c_void_p),
        ("features_act", ctypes.c_void_p),
        ("extra_data", ctypes.c_void_p), ("extra_rows", ctypes.c_int32), ("extra_cols", ctypes.c_int32),
        ("transformation_matrices", ctypes.c_void_p), ("features_act_stride", ctypes.c_int32),
        ("features_sigma", ctypes.c_void_p), ("accel_marks_current", ctypes.c_int32),
    ]


class _CTree64(ctypes.Structure):
    _fields_ = [
        ("features", ctypes.c_void_p), ("M", ctypes.c_int64), ("D", ctypes.c_int32), ("N", ctypes.c_int32),
        ("child", ctypes.c_void_p), ("data", ctypes.c_void_p), ("n_nodes", ctypes.c_int64), ("n_internal", ctypes.c_int64),
        ("offset", ctypes.c_void_p), ("scaling", ctypes.c_void_p),
    ]


class _CCamera64(ctypes.Structure):
    _fields_ = [("c2w", ctypes.c_void_p), ("fx", ctypes.c_double), ("fy", ctypes.c_double),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("row_begin", ctypes.c_int32), ("row_end", ctypes.c_int32)]


class _COptions(ctypes.Structure):
    _fields_ = [
        ("step_size", ctypes.c_float), ("background_brightness", ctypes.c_float),
        ("format", ctypes.c_int32), ("basis_dim", ctypes.c_int32),
        ("ndc_width", ctypes.c_int32), ("ndc_height", ctypes.c_int32), ("ndc_focal", ctypes.c_float),
        ("min_comp", ctypes.c_int32), ("max_comp", ctypes.c_int32),
        ("sigma_thresh", ctypes.c_float), ("stop_thresh", ctypes.c_float),
    ]


class _CPeerGroup(ctypes.Structure):
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("buffers", ctypes.POINTER(ctypes.c_void_p)),
                ("multicast", ctypes.c_void_p), ("table_offset", ctypes.c_int64), ("flags_offset", ctypes.c_int64),
                ("status_offset", ctypes.c_int64), ("blocks", ctypes.c_int32), ("epoch", ctypes.c_uint32)]


class _CCamera(ctypes.Structure):
    _fields_ = [("c2w", ctypes.c_void_p), ("fx", ctypes.c_float), ("fy", ctypes.c_float),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("row_begin", ctypes.c_int32), ("row_end", ctypes.c_int32)]


# Every symbol include/svoxb.h declares: (restype, argtypes). tests/test_cabi.py checks the list against the header.
_VP, _I64, _I32, _F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float
_PT, _PO, _PC = ctypes.POINTER(_CTree), ctypes.POINTER(_COptions), ctypes.POINTER(_CCamera)
_PT64, _PC64 = ctypes.POINTER(_CTree64), ctypes.POINTER(_CCamera64)
SYMBOLS = {
    "svoxb_abi_version": (ctypes.c_int, []),
    "svoxb_last_error": (ctypes.c_char_p, []),
    "svoxb_launch_count": (_I64, []),
    "svoxb_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3),
    "svoxb_accel_create": (ctypes.c_int, [_PT, ctypes.c_int, _VP, ctypes.POINTER(_VP)]),
    "svoxb_accel_rebuild": (ctypes.c_int, [_VP, _PT, ctypes.c_int, _VP]),
    "svoxb_accel_destroy": (None, [_VP]),
    "svoxb_accel_bytes": (_I64, [_VP]),
    "svoxb_accel_describe": (ctypes.c_int, [_VP, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(_I64)]),
    "svoxb_accel_mark_hits": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP]),
    "svoxb_activate_features": (ctypes.c_int, [_VP, _I64, _I32, _VP, _I32, _VP, _VP]),
    "svoxb_prepare_step": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP, _I32, _VP, _VP, _VP]),
    "svoxb_gather_sigma": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP]),
    "svoxb_query": (ctypes.c_int, [_PT, _VP, _I64, _VP, _VP, _VP, _VP, _VP]),
    "svoxb_leafset_scratch_bytes": (ctypes.c_size_t, [_I64]),
    "svoxb_leafset_scan": (ctypes.c_int, [_VP, _I64, _VP, _VP, _VP]),
    "svoxb_leafset_emit": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP, _VP]),
    "svoxb_construct_tree": (ctypes.c_int, [_PT, _VP, _VP, _I64, _VP]),
    "svoxb_query_bwd": (ctypes.c_int, [_PT, _VP, _I64, _VP, _I32, _VP, _VP]),
    "svoxb_assign": (ctypes.c_int, [_PT, _VP, _VP, _I64, _VP, _I32, _VP]),
    "svoxb_calc_corners": (ctypes.c_int, [_VP, _I32, _I64, _VP, _I64, _VP, _VP]),
    "svoxb_grid_weight_render": (ctypes.c_int, [_VP, _I32, _PC, _PO, _VP, _VP, _VP, _VP, _VP]),
    "svoxb_render_rays_fwd": (ctypes.c_int, [_PT, _VP, _VP, _VP, _I64, _PO, _VP, _VP, _VP]),
    "svoxb_out_data_dim": (ctypes.c_int, [_I32, _I32, _I32]),
    "svoxb_render_rays_bwd": (ctypes.c_int, [_PT, _VP, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_ray_order_max_rays": (_I64, []),
    "svoxb_ray_order_min_rays": (_I64, []),
    "svoxb_render_rays_fwd_cost": (ctypes.c_int, [_PT, _VP, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_render_rays_bwd_cost": (ctypes.c_int, [_PT, _VP, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _VP, _VP]),
    "svoxb_render_image_fwd": (ctypes.c_int, [_PT, _PC, _PO, _VP, _VP, _VP]),
    "svoxb_render_image_bwd": (ctypes.c_int, [_PT, _PC, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_render_depth": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP]),
    "svoxb_opacity_render_fwd": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP]),
    "svoxb_opacity_render_bwd": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP, _VP]),
    "svoxb_opacity_render_bwd_saved": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_motion_render": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _I32, _VP, _VP, _VP, _VP, _VP]),
    "svoxb_accumulate_weights": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PC, _PO, _VP, _VP]),
    "svoxb_motion_feature_render_fwd": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _I32, _I32, _I32, _VP, _VP]),
    "svoxb_motion_feature_render_bwd": (ctypes.c_int, [_PT, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _I32, _I32, _I32, _VP, _VP,
                                                       _VP]),
    "svoxb_warp_vertices": (ctypes.c_int, [_VP, _VP, _VP, _VP, _I64, _I32, _VP, _VP, _VP]),
    "svoxb_warp_vertices_bwd": (ctypes.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _I64, _I32, _I32, _VP, _VP, _VP, _VP]),
    "svoxb_p2v_bwd": (ctypes.c_int, [_VP, _VP, _VP, _I64, _I32, _VP, _VP, _I32, _F, _F, _VP, _VP, _VP]),
    "svoxb_p2v": (ctypes.c_int, [_VP, _VP, _I64, _I32, _VP, _VP, _I32, _F, _F, _VP, _VP]),
    "svoxb_exchange_max_blocks": (ctypes.c_int, []),
    "svoxb_exchange_sum": (ctypes.c_int, [ctypes.POINTER(_CPeerGroup), _I64, _VP]),
    "svoxb_exchange_sum_rows": (ctypes.c_int, [ctypes.POINTER(_CPeerGroup), _I64, _I32, _VP, _VP]),
    "svoxb_query_f64": (ctypes.c_int, [_PT64, _VP, _I64, _VP, _VP, _VP, _VP, _VP]),
    "svoxb_render_rays_fwd_f64": (ctypes.c_int, [_PT64, _VP, _VP, _I64, _PO, _VP, _VP, _VP]),
    "svoxb_render_rays_bwd_f64": (ctypes.c_int, [_PT64, _VP, _VP, _I64, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_render_image_fwd_f64": (ctypes.c_int, [_PT64, _PC64, _PO, _VP, _VP, _VP]),
    "svoxb_render_image_bwd_f64": (ctypes.c_int, [_PT64, _PC64, _PO, _VP, _VP, _VP, _VP]),
    "svoxb_render_depth_f64": (ctypes.c_int, [_PT64, _VP, _VP, _I64, _PO, _VP, _VP]),
    "svoxb_build_work_bytes": (ctypes.c_size_t, [_I64, _I32]),
    "svoxb_build_octree_count": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP, _VP, ctypes.POINTER(_I64), _VP]),
    "svoxb_build_octree_emit": (ctypes.c_int, [_I64, _I32, _VP, _I64, _VP, _VP, _VP, _VP]),
    "svoxb_build_dense_max_depth": (_I32, []),
    "svoxb_build_dense_work_bytes": (ctypes.c_size_t, [_I32]),
    "svoxb_build_dense": (ctypes.c_int, [_VP, _I64, _I32, _VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP]),
}

_lib = None


def load_library():
    """Load libsvoxb.so (built by ``__graft_entry__.build()`` / ``make -C svox_t_b200/csrc``). Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO_PATH):
            raise ImportError(
                f"svox_t_b200: the CUDA library {_SO_PATH} is missing and there is no CPU fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C svox_t_b200/csrc`.")
        lib = ctypes.CDLL(_SO_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.svoxb_abi_version() != 10:
            raise ImportError("svox_t_b200: libsvoxb.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError("svox_t_b200.csrc: " + load_library().svoxb_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load_library().svoxb_launch_count())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # the raw handle of torch's current stream on the current device (25 us cheaper per call than building a Stream object)
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def _check_input(t, name, dtype=None):
    # CHECK_INPUT of the reference (include/data_spec.hpp:38-43)
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype} (got {t.dtype}): features, offset, scaling, rays / points and "
                           "gradients must share one floating type (float32 everywhere; float64 for the point query, the "
                           "RGBA render forward / backward and depth)")


class _TensorIdentity:
    """What a derived table (activated features, hit marks, accelerator) was computed from: the tensor's STORAGE object,
    held strongly -- so its address cannot be handed to another tensor while this key is alive -- plus offset, shape,
    strides and the version counter. A fresh tensor that happens to reuse a freed block at the same address (every
    training step, with the caching allocator) is a different storage and never matches; views / ``detach()`` of the
    same tensor do. Writes through ``.data`` or a raw pointer bypass the version counter and cannot be seen: callers
    that do that must bump it (``torch.autograd.graph.increment_version``) -- the library's own in-place kernels do."""
    __slots__ = ("storage", "key")

    def __init__(self, t):
        self.storage = t.untyped_storage()
        self.key = self._key(t, self.storage)

    @staticmethod
    def _key(t, st):
        return (st._cdata, t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version)

    def matches(self, t):
        return t is not None and self._key(t, t.untyped_storage()) == self.key


# ---- spec records (field names of svox.cpp:74-117) --------------------------------------------------------------
class RaysSpec:
    def __init__(self):
        self.origins = None
        self.dirs = None
        self.vdirs = None
        self._cost = None           # svox_t_b200 extension: per-ray march cost written by the forward (short batches), the
                                    # backward's scheduling hint (svoxb_render_rays_fwd_cost / _bwd_cost)

    def check(self, dtype=torch.float32):
        for n in ("origins", "dirs", "vdirs"):
            _check_input(getattr(self, n), n, dtype)


class TreeSpec:
    def __init__(self):
        self.features = None
        self.data = None
        self.child = None
        self.parent_depth = None
        self.extra_data = None
        self.offset = None
        self.scaling = None
        self._weight_accum = None
        self.joint_features = None
        self.skinning_weights = None
        self.joint_index = None
        self.transformation_matrices = None
        self.n_internal = 0
        self._grad_exchange = None  # svox_t_b200 extension: dist.LeafGradExchange -- the backward reduces into its symmetric
                                    # table and sums it over the GPUs (None = a fresh zeros_like(features), as the reference)
        self._accel = None          # svox_t_b200 extension: Accel handle cached by N3Tree (None = reference walk)
        self._act = None            # svox_t_b200 extension: Activated table for `features` (None = sigmoid in-kernel)
        self._sigma = None          # svox_t_b200 extension: SigmaTable (compact sigma array) for the sigma-only marches

    @property
    def is_f64(self):
        """float64 instantiation (AT_DISPATCH_FLOATING_TYPES in the reference): decided by the feature table."""
        return isinstance(self.features, torch.Tensor) and self.features.dtype == torch.float64

    def _c64(self):
        """svoxb_tree_f64: the double-precision view of the tree (general kernels: no accelerator, no derived tables)."""
        _check_input(self.features, "features", torch.float64)
        _check_input(self.data, "data", torch.int32)
        _check_input(self.child, "child", torch.int32)
        _check_input(self.offset, "offset", torch.float64)
        _check_input(self.scaling, "scaling", torch.float64)
        if self.features.dim() != 2 or self.child.dim() != 4:
            raise RuntimeError("features must be [M, D] and child [n, N, N, N]")
        return _CTree64(features=_ptr(self.features), M=self.features.shape[0], D=self.features.shape[1],
                        N=self.child.shape[1], child=_ptr(self.child), data=_ptr(self.data),
                        n_nodes=self.child.shape[0], n_internal=int(self.n_internal),
                        offset=_ptr(self.offset), scaling=_ptr(self.scaling))

    def check(self):
        _check_input(self.features, "features", torch.float32)
        _check_input(self.data, "data", torch.int32)
        _check_input(self.child, "child", torch.int32)
        _check_input(self.parent_depth, "parent_depth", torch.int32)
        _check_input(self.offset, "offset", torch.float32)
        _check_input(self.scaling, "scaling", torch.float32)
        if self.features.dim() != 2 or self.child.dim() != 4:
            raise RuntimeError("features must be [M, D] and child [n, N, N, N]")
        if self._weight_accum is not None and self._weight_accum.numel():
            _check_input(self._weight_accum, "_weight_accum", torch.float32)
            if self._weight_accum.numel() != self.child.numel():
                raise RuntimeError("_weight_accum must have the shape of child")
        M = self.features.shape[0]
        ex, tm = self.extra_data, self.transformation_matrices
        if ex is not None and ex.numel():
            _check_input(ex, "extra_data", torch.float32)
            if ex.dim() != 2:
                raise RuntimeError("extra_data must be 2-D")
        if tm is not None and tm.numel():
            _check_input(tm, "transformation_matrices", torch.float32)
            if tuple(tm.shape) != (M, 4, 4):
                raise RuntimeError(f"transformation_matrices must be [M={M}, 4, 4]")

    def _c(self):
        self.check()
        acc = self._accel
        if acc is not None and not acc.matches(self):
            acc = None
        act = self._act
        if act is not None and not act.matches(self.features):
            act = None
        c = _CTree(
            features=_ptr(self.features), M=self.features.shape[0], D=self.features.shape[1],
            N=self.child.shape[1], child=_ptr(self.child), data=_ptr(self.data),
            parent_depth=_ptr(self.parent_depth), n_nodes=self.child.shape[0], n_internal=int(self.n_internal),
            offset=_ptr(self.offset), scaling=_ptr(self.scaling),
            accel=acc.handle if acc is not None else ctypes.c_void_p(0),
            features_act=_ptr(act.table) if act is not None else ctypes.c_void_p(0),
            extra_data=_ptr(self.extra_data),
            extra_rows=self.extra_data.shape[0] if self.extra_data is not None and self.extra_data.numel() else 0,
            extra_cols=self.extra_data.shape[1] if self.extra_data is not None and self.extra_data.numel() else 0,
            transformation_matrices=_ptr(self.transformation_matrices),
            features_act_stride=act.table.shape[1] if act is not None else 0,
            features_sigma=_ptr(act.sigma) if act is not None else (
                _ptr(self._sigma.table) if self._sigma is not None and self._sigma.matches(self.features)
                else ctypes.c_void_p(0)),
            accel_marks_current=1 if acc is not None and acc.marks_match(self.features) else 0)
        return c


class CameraSpec:
    def __init__(self):
        self.c2w = None
        self.fx = 0.0
        self.fy = 0.0
        self.width = 0
        self.height = 0
        self.row_begin = 0          # svox_t_b200 extension: render only rows [row_begin, row_end) (0, 0 = all)
        self.row_end = 0

    def check(self):
        _check_input(self.c2w, "c2w", torch.float32)
        if self.c2w.dim() != 2 or self.c2w.shape[1] != 4 or self.c2w.shape[0] < 3:
            raise RuntimeError("c2w must be [3 or 4, 4]")

    def _c64(self):
        _check_input(self.c2w, "c2w", torch.float64)
        if self.c2w.dim() != 2 or self.c2w.shape[1] != 4 or self.c2w.shape[0] < 3:
            raise RuntimeError("c2w must be [3 or 4, 4]")
        return _CCamera64(c2w=_ptr(self.c2w), fx=float(self.fx), fy=float(self.fy), width=int(self.width),
                          height=int(self.height), row_begin=int(self.row_begin), row_end=int(self.row_end))

    def _c(self):
        self.check()
        return _CCamera(c2w=_ptr(self.c2w), fx=float(self.fx), fy=float(self.fy), width=int(self.width),
                        height=int(self.height), row_begin=int(self.row_begin), row_end=int(self.row_end))

    @property
    def rows(self):
        """Number of image rows the kernels write (the band, or the whole image)."""
        return int(self.row_end - self.row_begin) if self.row_end > 0 else int(self.height)


class RenderOptions:
    def __init__(self):
        self.step_size = 1e-3
        self.background_brightness = 1.0
        self.format = FORMAT_RGBA
        self.basis_dim = -1
        self.ndc_width = -1
        self.ndc_height = -1
        self.ndc_focal = 0.0
        self.min_comp = 0
        self.max_comp = -1
        self.sigma_thresh = 0.0
        self.stop_thresh = 0.0

    def _c(self, **override):
        v = dict(step_size=self.step_size, background_brightness=self.background_brightness,
                 format=int(self.format), basis_dim=int(self.basis_dim), ndc_width=int(self.ndc_width),
                 ndc_height=int(self.ndc_height), ndc_focal=float(self.ndc_focal), min_comp=int(self.min_comp),
                 max_comp=int(self.max_comp), sigma_thresh=self.sigma_thresh, stop_thresh=self.stop_thresh)
        v.update(override)
        return _COptions(**v)


class Accel:
    """Owner of a packed grid+brick accelerator (svoxb_accel_create). Valid while child/data are unchanged."""

    def __init__(self, tree_spec, max_depth=0):
        lib = load_library()
        tree_spec._accel = None
        self._key = self._make_key(tree_spec)
        h = ctypes.c_void_p(0)
        with torch.cuda.device(tree_spec.child.device):
            _check(lib.svoxb_accel_create(ctypes.byref(tree_spec._c()), int(max_depth), _stream(), ctypes.byref(h)))
        self.handle = h
        self._lib = lib

    @staticmethod
    def _make_key(ts):
        return (_TensorIdentity(ts.child), _TensorIdentity(ts.data), int(ts.features.shape[0]), int(ts.n_internal))

    def matches(self, ts):
        k = self._key
        return bool(self.handle) and k[0].matches(ts.child) and k[1].matches(ts.data) and \
            k[2:] == (int(ts.features.shape[0]), int(ts.n_internal))

    def rebuild(self, tree_spec, max_depth):
        """Refill this accelerator for another child/data of the same depth, in place (svoxb_accel_rebuild: no
        allocation, no host synchronisation -- the per-frame path). False if it has to be created anew."""
        if not self.handle:
            return False
        tree_spec._accel = None
        with torch.cuda.device(tree_spec.child.device):
            rc = self._lib.svoxb_accel_rebuild(self.handle, ctypes.byref(tree_spec._c()), int(max_depth), _stream())
        if rc != 0:
            return False
        self._key = self._make_key(tree_spec)
        self._marks_key = None
        return True

    def mark_hits(self, features):
        """Refresh the per-leaf "sigma <= 0" marks for this exact (storage, version) of ``features``; the march then
        never fetches those rows. No-op when the marks are current."""
        if self.marks_match(features):
            return
        _check_input(features, "features", torch.float32)
        with torch.cuda.device(features.device):
            _check(self._lib.svoxb_accel_mark_hits(self.handle, _ptr(features), features.shape[0], features.shape[1],
                                                   _stream()))
        self._marks_key = _TensorIdentity(features)

    def marks_match(self, features):
        k = getattr(self, "_marks_key", None)
        return k is not None and k.matches(features)

    @property
    def nbytes(self):
        return int(self._lib.svoxb_accel_bytes(self.handle))

    def describe(self):
        n = ctypes.c_int(0)
        bits = (ctypes.c_int * 4)()
        bricks = (ctypes.c_int64 * 4)()
        _check(self._lib.svoxb_accel_describe(self.handle, ctypes.byref(n), bits, bricks))
        return dict(stages=n.value, bits=list(bits)[:n.value], bricks=list(bricks)[:n.value], bytes=self.nbytes)

    def __del__(self):
        try:
            if self.handle:
                self._lib.svoxb_accel_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Activated:
    """features with the sigmoid applied once per row to the payload channels (svoxb_prepare_step). Valid for
    exactly one (storage object, version, layout) of ``features`` (_TensorIdentity); the renderer rebuilds it whenever
    features change or another tensor is passed -- also one that reuses the address of a freed one.

    The same pass over the rows can do the other per-step table work: ``accel`` -- refresh that accelerator's hit marks
    for these features; ``zero_table`` -- zero-fill the [M, D] gradient table the step's backward will reduce into."""

    def __init__(self, features, accel=None, zero_table=None):
        lib = load_library()
        _check_input(features, "features", torch.float32)
        self._key = _TensorIdentity(features)
        M, D = features.shape
        if zero_table is not None:
            _check_input(zero_table, "zero_table", torch.float32)
            if tuple(zero_table.shape) != (M, D):
                raise RuntimeError("zero_table must have the shape of features")
        with torch.cuda.device(features.device):
            if D % 4 == 0:
                self.table, self.sigma = torch.empty_like(features), None
                stride = D
            else:       # payload-only aligned rows + compact sigma: every width runs on the 128-bit row kernels
                stride = (D - 1 + 3) // 4 * 4
                self.table = torch.empty((M, stride), dtype=torch.float32, device=features.device)
                self.sigma = torch.empty((M,), dtype=torch.float32, device=features.device)
            _check(lib.svoxb_prepare_step(accel.handle if accel is not None else None, _ptr(features), M, D,
                                          _ptr(self.table), stride, _ptr(self.sigma), _ptr(zero_table), _stream()))
        if accel is not None:
            accel._marks_key = _TensorIdentity(features)

    def matches(self, features):
        return self._key.matches(features)


class SigmaTable:
    """The sigma channel of ``features`` as a compact [M] array (svoxb_gather_sigma), for the marches that read nothing
    else of a row (depth, opacity, motion). Valid for one (storage object, version, layout) of ``features``."""

    def __init__(self, features):
        lib = load_library()
        _check_input(features, "features", torch.float32)
        self._key = _TensorIdentity(features)
        M, D = features.shape
        with torch.cuda.device(features.device):
            self.table = torch.empty((M,), dtype=torch.float32, device=features.device)
            _check(lib.svoxb_gather_sigma(_ptr(features), M, D, _ptr(self.table), _stream()))

    def matches(self, features):
        return self._key.matches(features)


# ---- functions (svox.cpp:119-144) -----------------------------------------------------------------------------------
def query_vertical(tree, indices):
    """(values[Q,D], node_ids[Q] i64, data_ids[Q] i64, leaf_node[n_hit,4] i64) -- svox_kernel.cu:274-324.
    Rows of ``values`` / ``data_ids`` whose leaf is empty are zero / -1 here (uninitialised in the reference)."""
    lib = load_library()
    f64 = tree.is_f64
    real = torch.float64 if f64 else torch.float32
    _check_input(indices, "indices", real)
    if indices.dim() != 2 or indices.shape[1] != 3:
        raise RuntimeError("indices must be [Q, 3]")
    ct = tree._c64() if f64 else tree._c()
    dev = indices.device
    Q, D = indices.shape[0], tree.features.shape[1]
    N = tree.child.shape[1]
    with torch.cuda.device(dev):
        values = torch.zeros((Q, D), dtype=real, device=dev)
        node_ids = torch.empty((Q,), dtype=torch.int64, device=dev)
        data_ids = torch.full((Q,), -1, dtype=torch.int64, device=dev)
        n_slots = int(tree.n_internal) * N ** 3
        mask = torch.zeros((n_slots,), dtype=torch.uint8, device=dev)
        _check((lib.svoxb_query_f64 if f64 else lib.svoxb_query)(
            ctypes.byref(ct), _ptr(indices), Q, _ptr(values), _ptr(node_ids), _ptr(data_ids), _ptr(mask), _stream()))
        scratch = torch.empty((lib.svoxb_leafset_scratch_bytes(n_slots),), dtype=torch.uint8, device=dev)
        n_hit = torch.zeros((1,), dtype=torch.int64, device=dev)
        _check(lib.svoxb_leafset_scan(_ptr(mask), n_slots, _ptr(scratch), _ptr(n_hit), _stream()))
        n = int(n_hit.item())                         # the one sync the reference also has (svox_kernel.cu:312)
        leaf_node = torch.empty((n, 4), dtype=torch.int64, device=dev)
        if n:
            _check(lib.svoxb_leafset_emit(_ptr(mask), n_slots, N, _ptr(scratch), _ptr(leaf_node), _stream()))
    return values, node_ids, data_ids, leaf_node


def construct_tree(tree, indices):
    """data[leaf(p_i)] = i, in place (svox_kernel.cu:341-352). Deterministic: the largest i wins a shared leaf."""
    lib = load_library()
    _check_input(indices, "indices", torch.float32)
    ct = tree._c()
    with torch.cuda.device(indices.device):
        _check(lib.svoxb_construct_tree(ctypes.byref(ct), _ptr(tree.data), _ptr(indices), indices.shape[0], _stream()))
    tree.data.add_(0)   # bump the tensor version: cached accelerators built from the old data are now stale


def _out_dim(tree, opt):
    """Output row width (get_out_data_dim, rt_kernel.cu:1352-1358)."""
    n = load_library().svoxb_out_data_dim(int(opt.format), int(opt.basis_dim), tree.features.shape[1])
    if n <= 0:
        raise RuntimeError(f"svox_t_b200.csrc: bad format/basis_dim ({opt.format}, {opt.basis_dim})")
    return n


def _accumulate_weights(tree, ct, rays, cam_c, opt):
    """TreeSpec._weight_accum side effect of the two render entry points (rt_kernel.cu:308-310)."""
    wa = tree._weight_accum
    if wa is None or not wa.numel():
        return
    lib = load_library()
    o, d, Q = (_ptr(rays.origins), _ptr(rays.dirs), rays.origins.shape[0]) if rays is not None else (None, None, 0)
    _check(lib.svoxb_accumulate_weights(ctypes.byref(ct), o, d, Q, ctypes.byref(cam_c) if cam_c is not None else None,
                                        ctypes.byref(opt._c()), _ptr(wa), _stream()))


_ORDER_RANGE = None


def _order_range():
    """Batch sizes for which the forward leaves the backward its per-ray march costs (svoxb_order.cu)."""
    global _ORDER_RANGE
    if _ORDER_RANGE is None:
        lib = load_library()
        _ORDER_RANGE = (int(lib.svoxb_ray_order_min_rays()), int(lib.svoxb_ray_order_max_rays()))
    return _ORDER_RANGE


def _require_rgba_f64(opt, tree=None):
    if int(opt.format) != FORMAT_RGBA:
        raise RuntimeError("svox_t_b200.csrc: float64 is implemented for the RGBA format only (view-dependent formats "
                           "are float32)")
    wa = getattr(tree, "_weight_accum", None)
    if wa is not None and wa.numel():
        raise RuntimeError("svox_t_b200.csrc: accumulate_weights is float32 only")


def _render_fwd_f64(tree, rays, opt, want_depth, c_opt=None):
    lib = load_library()
    _require_rgba_f64(opt, tree)
    rays.check(torch.float64)
    ct = tree._c64()
    Q, D = rays.origins.shape[0], tree.features.shape[1]
    dev = rays.origins.device
    with torch.cuda.device(dev):
        out = torch.empty((Q, D), dtype=torch.float64, device=dev)
        depth = torch.empty((Q, 1), dtype=torch.float64, device=dev) if want_depth else None
        _check(lib.svoxb_render_rays_fwd_f64(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                             ctypes.byref(c_opt if c_opt is not None else opt._c()), _ptr(out),
                                             _ptr(depth), _stream()))
    return out, depth


def _render_fwd(tree, rays, opt, want_depth):
    if tree.is_f64:
        return _render_fwd_f64(tree, rays, opt, want_depth)
    lib = load_library()
    rays.check()
    ct = tree._c()
    Q, D = rays.origins.shape[0], _out_dim(tree, opt)
    dev = rays.origins.device
    with torch.cuda.device(dev):
        out = torch.empty((Q, D), dtype=torch.float32, device=dev)
        depth = None
        if want_depth and opt.format != FORMAT_RGBA:       # the view-dependent kernels have no fused depth output
            depth = torch.empty((Q, 1), dtype=torch.float32, device=dev)
            _check(lib.svoxb_render_depth(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                          ctypes.byref(opt._c()), _ptr(depth), _stream()))
            want_depth = False
        fused = torch.empty((Q, 1), dtype=torch.float32, device=dev) if want_depth else None
        depth = fused if want_depth else depth
        rays._cost = None
        lo, hi = _order_range()
        if ct.accel and lo <= Q <= hi:                      # short batch: the forward leaves the backward its completion list
            rays._cost = torch.empty((Q,), dtype=torch.int32, device=dev)
        _check(lib.svoxb_render_rays_fwd_cost(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), _ptr(rays.vdirs), Q,
                                              ctypes.byref(opt._c()), _ptr(out), _ptr(fused), _ptr(rays._cost), _stream()))
        _accumulate_weights(tree, ct, rays, None, opt)
    return out, depth


def volume_render(tree, rays, opt):
    """[Q, D] composited features + opacity (rt_kernel.cu:1362-1379)."""
    return _render_fwd(tree, rays, opt, False)[0]


def volume_render_with_depth(tree, rays, opt):
    """svox_t_b200 extension: (out[Q,D], depth[Q,1]) from ONE march (the reference needs a second kernel)."""
    return _render_fwd(tree, rays, opt, True)


def _saved_out_for_backward(tree, opt, fwd_out, render_again, expect_shape=None):
    # The backward's hit predicate is sigma > 0 with no early stop (rt_kernel.cu:382,456). With the default
    # options the forward output IS the backward's saved state; otherwise re-render with those semantics.
    # ``fwd_out`` must be the UNMODIFIED output of the forward for the same tree, features, rays and options: the
    # backward reads T_end = 1 - out[:, D-1] and <grad_out, out> from it. Anything else (None, another shape, another
    # dtype / device) is not trusted and the state is re-rendered.
    if (fwd_out is not None and expect_shape is not None and tuple(fwd_out.shape) == tuple(expect_shape)
            and fwd_out.dtype == tree.features.dtype and fwd_out.is_cuda and fwd_out.is_contiguous()
            and opt.sigma_thresh == 0.0 and opt.stop_thresh <= 0.0):
        return fwd_out
    return render_again()


def volume_render_backward(tree, rays, opt, grad_output, saved_out=None):
    """[M, D] dL/dfeatures (rt_kernel.cu:1402-1426). ``saved_out`` = the forward output, if the caller kept it."""
    lib = load_library()
    if tree.is_f64:
        _require_rgba_f64(opt)
        if getattr(tree, "_grad_exchange", None) is not None:
            raise RuntimeError("svox_t_b200.csrc: the multi-GPU leaf-gradient exchange is float32 only")
        rays.check(torch.float64)
        _check_input(grad_output, "grad_output", torch.float64)
        ct = tree._c64()
        Q = rays.origins.shape[0]
        bopt = opt._c(sigma_thresh=0.0, stop_thresh=-1.0)
        with torch.cuda.device(rays.origins.device):
            so = _saved_out_for_backward(tree, opt, saved_out, lambda: _render_fwd_f64(tree, rays, opt, False, bopt)[0],
                                         grad_output.shape)
            grad = torch.zeros_like(tree.features)
            _check(lib.svoxb_render_rays_bwd_f64(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                                 ctypes.byref(bopt), _ptr(grad_output), _ptr(so), _ptr(grad), _stream()))
        return grad
    rays.check()
    _check_input(grad_output, "grad_output", torch.float32)
    ct = tree._c()
    Q = rays.origins.shape[0]
    dev = rays.origins.device
    bopt = opt._c(sigma_thresh=0.0, stop_thresh=-1.0)

    def again():
        o = torch.empty_like(grad_output)
        _check(lib.svoxb_render_rays_fwd(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), _ptr(rays.vdirs), Q,
                                         ctypes.byref(bopt), _ptr(o), ctypes.c_void_p(0), _stream()))
        return o

    with torch.cuda.device(dev):
        so = _saved_out_for_backward(tree, opt, saved_out, again, grad_output.shape)
        xchg = getattr(tree, "_grad_exchange", None)
        if xchg is not None:
            # The gradient handed to autograd IS the exchange's table (no 243 MB copy). If features.grad still aliases it
            # from an earlier backward (gradient accumulation over several backward calls), give that gradient a life
            # of its own before the table is zeroed and reused -- autograd then adds the new sum to the copy.
            fg = getattr(tree.features, "grad", None) if tree.features.is_leaf else None
            if fg is not None and fg.data_ptr() == xchg.table.data_ptr():
                tree.features.grad = fg.clone()
        # multi-GPU: reduce into the exchange's symmetric table and sum it over the ranks right here (dist.LeafGradExchange)
        grad = xchg.table_for_backward(tree.features) if xchg is not None else torch.zeros_like(tree.features)
        cost = getattr(rays, "_cost", None)
        if cost is not None and (cost.shape[0] != Q or cost.device != dev):
            cost = None
        _check(lib.svoxb_render_rays_bwd_cost(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), _ptr(rays.vdirs), Q,
                                              ctypes.byref(bopt), _ptr(grad_output), _ptr(so), _ptr(grad), _ptr(cost),
                                              _stream()))
        if xchg is not None:
            xchg.all_reduce_(features=tree.features)     # rows with sigma <= 0 hold zeros on every rank: skipped
            grad = grad.view(grad.shape)        # a fresh tensor object over the table: autograd adopts it without a copy
    return grad


def _render_image_fwd_f64(tree, cam, opt, want_depth, c_opt=None):
    lib = load_library()
    _require_rgba_f64(opt, tree)
    ct, cc = tree._c64(), cam._c64()
    dev = tree.features.device
    D = tree.features.shape[1]
    with torch.cuda.device(dev):
        out = torch.empty((cam.rows, cam.width, D), dtype=torch.float64, device=dev)
        depth = torch.empty((cam.rows, cam.width, 1), dtype=torch.float64, device=dev) if want_depth else None
        _check(lib.svoxb_render_image_fwd_f64(ctypes.byref(ct), ctypes.byref(cc),
                                              ctypes.byref(c_opt if c_opt is not None else opt._c()), _ptr(out),
                                              _ptr(depth), _stream()))
    return out, depth


def _render_image_fwd(tree, cam, opt, want_depth):
    if tree.is_f64:
        return _render_image_fwd_f64(tree, cam, opt, want_depth)
    lib = load_library()
    ct, cc = tree._c(), cam._c()
    dev = tree.features.device
    D = _out_dim(tree, opt)
    if want_depth and opt.format != FORMAT_RGBA:
        raise RuntimeError("svox_t_b200.csrc: fused image depth is only available for the RGBA format")
    with torch.cuda.device(dev):
        out = torch.empty((cam.rows, cam.width, D), dtype=torch.float32, device=dev)
        depth = torch.empty((cam.rows, cam.width, 1), dtype=torch.float32, device=dev) if want_depth else None
        _check(lib.svoxb_render_image_fwd(ctypes.byref(ct), ctypes.byref(cc), ctypes.byref(opt._c()), _ptr(out),
                                          _ptr(depth), _stream()))
        _accumulate_weights(tree, ct, None, cc, opt)
    return out, depth


def volume_render_image(tree, cam, opt):
    """[H, W, D] (rt_kernel.cu:1381-1400; broken in the reference by its int32 dtype dispatch)."""
    return _render_image_fwd(tree, cam, opt, False)[0]


def volume_render_image_with_depth(tree, cam, opt):
    return _render_image_fwd(tree, cam, opt, True)


def volume_render_image_backward(tree, cam, opt, grad_output, saved_out=None):
    lib = load_library()
    if tree.is_f64:
        _require_rgba_f64(opt)
        _check_input(grad_output, "grad_output", torch.float64)
        ct, cc = tree._c64(), cam._c64()
        bopt = opt._c(sigma_thresh=0.0, stop_thresh=-1.0)
        with torch.cuda.device(tree.features.device):
            so = _saved_out_for_backward(tree, opt, saved_out,
                                         lambda: _render_image_fwd_f64(tree, cam, opt, False, bopt)[0], grad_output.shape)
            grad = torch.zeros_like(tree.features)
            _check(lib.svoxb_render_image_bwd_f64(ctypes.byref(ct), ctypes.byref(cc), ctypes.byref(bopt),
                                                  _ptr(grad_output), _ptr(so), _ptr(grad), _stream()))
        return grad
    _check_input(grad_output, "grad_output", torch.float32)
    ct, cc = tree._c(), cam._c()
    dev = tree.features.device
    bopt = opt._c(sigma_thresh=0.0, stop_thresh=-1.0)

    def again():
        o = torch.empty_like(grad_output)
        _check(lib.svoxb_render_image_fwd(ctypes.byref(ct), ctypes.byref(cc), ctypes.byref(bopt), _ptr(o),
                                          ctypes.c_void_p(0), _stream()))
        return o

    with torch.cuda.device(dev):
        so = _saved_out_for_backward(tree, opt, saved_out, again, grad_output.shape)
        grad = torch.zeros_like(tree.features)
        _check(lib.svoxb_render_image_bwd(ctypes.byref(ct), ctypes.byref(cc), ctypes.byref(bopt), _ptr(grad_output),
                                          _ptr(so), _ptr(grad), _stream()))
    return grad


def render_depth(tree, rays, opt):
    """[Q, 1] first-hit depth (rt_kernel.cu:1506-1523)."""
    lib = load_library()
    if tree.is_f64:
        rays.check(torch.float64)
        ct = tree._c64()
        Q = rays.origins.shape[0]
        with torch.cuda.device(rays.origins.device):
            depth = torch.empty((Q, 1), dtype=torch.float64, device=rays.origins.device)
            _check(lib.svoxb_render_depth_f64(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                              ctypes.byref(opt._c()), _ptr(depth), _stream()))
        return depth
    rays.check()
    ct = tree._c()
    Q = rays.origins.shape[0]
    dev = rays.origins.device
    with torch.cuda.device(dev):
        depth = torch.empty((Q, 1), dtype=torch.float32, device=dev)
        _check(lib.svoxb_render_depth(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                      ctypes.byref(opt._c()), _ptr(depth), _stream()))
    return depth


def opacity_render(tree, rays, opt):
    """[Q, 1] opacity 1 - T (rt_kernel.cu:1574-1591)."""
    lib = load_library()
    rays.check()
    ct = tree._c()
    Q, dev = rays.origins.shape[0], rays.origins.device
    with torch.cuda.device(dev):
        out = torch.empty((Q, 1), dtype=torch.float32, device=dev)
        _check(lib.svoxb_opacity_render_fwd(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                            ctypes.byref(opt._c()), _ptr(out), _stream()))
    return out


def opacity_render_backward(tree, rays, opt, grad_output, saved_out=None):
    """[M, D] gradient of opacity_render w.r.t. the sigma channel -- the semantics of the reference's
    opacity_trace_ray_backward (rt_kernel.cu:562-651), which its own wrapper never launches (Appendix B2).
    ``saved_out``: the UNMODIFIED output of opacity_render for the same tree, rays and default thresholds; the backward
    then marches once (T_end = 1 - saved_out) instead of twice."""
    lib = load_library()
    rays.check()
    _check_input(grad_output, "grad_output", torch.float32)
    ct = tree._c()
    Q, dev = rays.origins.shape[0], rays.origins.device
    use_saved = (saved_out is not None and tuple(saved_out.shape) == (Q, 1) and saved_out.dtype == torch.float32
                 and saved_out.is_cuda and saved_out.is_contiguous() and opt.sigma_thresh == 0.0 and opt.stop_thresh <= 0.0)
    with torch.cuda.device(dev):
        grad = torch.zeros_like(tree.features)
        if use_saved:
            _check(lib.svoxb_opacity_render_bwd_saved(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                                      ctypes.byref(opt._c()), _ptr(grad_output), _ptr(saved_out), _ptr(grad),
                                                      _stream()))
        else:
            _check(lib.svoxb_opacity_render_bwd(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                                ctypes.byref(opt._c()), _ptr(grad_output), _ptr(grad), _stream()))
    return grad


def motion_render(tree, rays, opt):
    """[out[Q,J], depth[Q,1], hit_point[Q,3], data_idx[Q,1] i64] at the first hit (rt_kernel.cu:1480-1504)."""
    lib = load_library()
    rays.check()
    if tree.extra_data is None or tree.extra_data.numel() == 0:
        raise RuntimeError("motion_render needs tree.extra_data [J, 3]")
    _check_input(tree.extra_data, "extra_data", torch.float32)
    ct = tree._c()
    Q, dev, J = rays.origins.shape[0], rays.origins.device, tree.extra_data.shape[0]
    with torch.cuda.device(dev):
        out = torch.empty((Q, J), dtype=torch.float32, device=dev)
        depth = torch.empty((Q, 1), dtype=torch.float32, device=dev)
        hit = torch.empty((Q, 3), dtype=torch.float32, device=dev)
        didx = torch.empty((Q, 1), dtype=torch.int64, device=dev)
        _check(lib.svoxb_motion_render(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q, ctypes.byref(opt._c()),
                                       _ptr(tree.extra_data), J, _ptr(out), _ptr(depth), _ptr(hit), _ptr(didx),
                                       _stream()))
    return [out, depth, hit, didx]


def _joint_args(tree):
    jf, sw, ji = tree.joint_features, tree.skinning_weights, tree.joint_index
    if jf is None or sw is None or ji is None or not jf.numel():
        raise RuntimeError("motion_feature_render needs tree.joint_features [J,F], skinning_weights [M,B], joint_index [M,B]")
    _check_input(jf, "joint_features", torch.float32)
    _check_input(sw, "skinning_weights", torch.float32)
    _check_input(ji, "joint_index", torch.int32)
    M = tree.features.shape[0]
    if jf.dim() != 2 or sw.dim() != 2 or tuple(ji.shape) != tuple(sw.shape) or sw.shape[0] != M:
        raise RuntimeError(f"joint_features must be [J,F], skinning_weights / joint_index [M={M},B]")
    return jf, sw, ji


def motion_feature_render(tree, rays, opt):
    """[Q, F] composited per-joint features (rt_kernel.cu:1525-1543)."""
    lib = load_library()
    rays.check()
    jf, sw, ji = _joint_args(tree)
    ct = tree._c()
    Q, dev = rays.origins.shape[0], rays.origins.device
    with torch.cuda.device(dev):
        out = torch.empty((Q, jf.shape[1]), dtype=torch.float32, device=dev)
        _check(lib.svoxb_motion_feature_render_fwd(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                                   ctypes.byref(opt._c()), _ptr(jf), _ptr(sw), _ptr(ji), jf.shape[0],
                                                   jf.shape[1], sw.shape[1], _ptr(out), _stream()))
    return out


def motion_feature_render_backward(tree, rays, opt, grad_output):
    """[J, F] dL/d joint_features: the gradient the reference's kernel set out to compute (rt_kernel.cu:981-1064 adds
    into an uninitialised array and indexes it by bone slot)."""
    lib = load_library()
    rays.check()
    _check_input(grad_output, "grad_output", torch.float32)
    jf, sw, ji = _joint_args(tree)
    ct = tree._c()
    Q, dev = rays.origins.shape[0], rays.origins.device
    with torch.cuda.device(dev):
        grad = torch.empty_like(jf)
        _check(lib.svoxb_motion_feature_render_bwd(ctypes.byref(ct), _ptr(rays.origins), _ptr(rays.dirs), Q,
                                                   ctypes.byref(opt._c()), _ptr(jf), _ptr(sw), _ptr(ji), jf.shape[0],
                                                   jf.shape[1], sw.shape[1], _ptr(grad_output), _ptr(grad), _stream()))
    return grad


def warp_vertices(matrices, indices, skinning_weights, joint_index):
    """[coords'[P,3], mats[P,4,4]] (svox_kernel.cu:354-378)."""
    lib = load_library()
    _check_input(indices, "indices", torch.float32)
    _check_input(matrices, "matrices", torch.float32)
    _check_input(skinning_weights, "skinning_weights", torch.float32)
    _check_input(joint_index, "joint_index", torch.int32)
    P, B = skinning_weights.shape
    dev = indices.device
    with torch.cuda.device(dev):
        vout = torch.empty((P, 3), dtype=torch.float32, device=dev)
        mout = torch.empty((P, 4, 4), dtype=torch.float32, device=dev)
        _check(lib.svoxb_warp_vertices(_ptr(matrices), _ptr(indices), _ptr(skinning_weights), _ptr(joint_index), P, B,
                                       _ptr(vout), _ptr(mout), _stream()))
    return [vout, mout]


def warp_vertices_backward(matrices, indices, skinning_weights, joint_index, indices_grad_out, matrices_grad_out):
    """[grad_indices[P,3], grad_matrices[J,4,4], grad_skinning_weights[P,B]] (svox_kernel.cu:404-436)."""
    lib = load_library()
    for t, n in ((matrices, "matrices"), (indices, "indices"), (skinning_weights, "skinning_weights"),
                 (indices_grad_out, "indices_grad_out"), (matrices_grad_out, "matrices_grad_out")):
        _check_input(t, n, torch.float32)
    _check_input(joint_index, "joint_index", torch.int32)
    P, B = skinning_weights.shape
    J, dev = matrices.shape[0], indices.device
    with torch.cuda.device(dev):
        g_T = torch.empty((J, 4, 4), dtype=torch.float32, device=dev)
        g_x = torch.empty((P, 3), dtype=torch.float32, device=dev)
        g_w = torch.empty((P, B), dtype=torch.float32, device=dev)
        _check(lib.svoxb_warp_vertices_bwd(_ptr(matrices), _ptr(indices), _ptr(skinning_weights), _ptr(joint_index),
                                           _ptr(indices_grad_out), _ptr(matrices_grad_out), P, B, J, _ptr(g_T), _ptr(g_x),
                                           _ptr(g_w), _stream()))
    return [g_x, g_T, g_w]


def p2v_backward(grad_output, points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
    """[points_grad[P,3], point_features_grad[P,F]] (p2v_kernel.cu:263-285)."""
    lib = load_library()
    for t, n in ((grad_output, "grad_output"), (points, "points"), (point_features, "point_features"),
                 (volume_corner, "volume_corner"), (volume_size, "volume_size")):
        _check_input(t, n, torch.float32)
    P, F, dev = points.shape[0], point_features.shape[1], points.device
    with torch.cuda.device(dev):
        g_p = torch.empty((P, 3), dtype=torch.float32, device=dev)
        g_f = torch.empty((P, F), dtype=torch.float32, device=dev)
        _check(lib.svoxb_p2v_bwd(_ptr(grad_output), _ptr(points), _ptr(point_features), P, F, _ptr(volume_corner),
                                 _ptr(volume_size), int(n_voxels), float(kernel_radius), float(conv_radius), _ptr(g_p),
                                 _ptr(g_f), _stream()))
    return [g_p, g_f]


def p2v(points, point_features, volume_corner, volume_size, n_voxels, kernel_radius, conv_radius):
    """[n, n, n, 1] Gaussian splat (p2v_kernel.cu:240-261)."""
    lib = load_library()
    for t, n in ((points, "points"), (point_features, "point_features"), (volume_corner, "volume_corner"),
                 (volume_size, "volume_size")):
        _check_input(t, n, torch.float32)
    dev = points.device
    with torch.cuda.device(dev):
        vox = torch.empty((n_voxels, n_voxels, n_voxels, 1), dtype=torch.float32, device=dev)
        _check(lib.svoxb_p2v(_ptr(points), _ptr(point_features), points.shape[0], point_features.shape[1],
                             _ptr(volume_corner), _ptr(volume_size), int(n_voxels), float(kernel_radius),
                             float(conv_radius), _ptr(vox), _stream()))
    return vox


_BUILD_WORK = {}


def build_octree(points, depth, offset, scaling, capacity=None, sort_based=False):
    """One-shot octree of finest level ``depth`` whose leaves at that level are the cells occupied by ``points``
    (svox_t_b200 extension; replaces depth-1 rounds of tree[pts].refine() + construct_tree, svox.py:488-560).
    Returns (child[n,2,2,2], data[n,2,2,2,1], parent_depth[n,2], status) int32, reference format; data = point index.

    ``capacity=None``: tensors of exactly the nodes needed (one host read-back of the node count), status None.
    ``capacity=n``: tensors of n nodes, NO host synchronisation; ``status`` is a device int64[2] = (nodes needed,
    overflow flag) to be read whenever convenient -- the per-frame path. Depths beyond the bitmap build's limit (10),
    or ``sort_based=True``, take the sort-based build (svoxb_build.cu)."""
    lib = load_library()
    _check_input(points, "points", torch.float32)
    _check_input(offset, "offset", torch.float32)
    _check_input(scaling, "scaling", torch.float32)
    P, dev, depth = points.shape[0], points.device, int(depth)
    with torch.cuda.device(dev):
        if sort_based or depth > lib.svoxb_build_dense_max_depth():
            if capacity is not None:
                raise RuntimeError("the sort-based build returns exactly-sized tensors (capacity must be None)")
            nbytes = lib.svoxb_build_work_bytes(P, depth)
            if nbytes == 0:
                raise RuntimeError("svox_t_b200.csrc.build_octree: depth/point count out of range")
            work = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
            n = ctypes.c_int64(0)
            _check(lib.svoxb_build_octree_count(_ptr(points), P, depth, _ptr(offset), _ptr(scaling), _ptr(work),
                                                ctypes.byref(n), _stream()))
            n = int(n.value)
            child = torch.empty((n, 2, 2, 2), dtype=torch.int32, device=dev)
            data = torch.empty((n, 2, 2, 2, 1), dtype=torch.int32, device=dev)
            parent_depth = torch.empty((n, 2), dtype=torch.int32, device=dev)
            _check(lib.svoxb_build_octree_emit(P, depth, _ptr(work), n, _ptr(child), _ptr(data), _ptr(parent_depth),
                                               _stream()))
            return child, data, parent_depth, None
        key = (dev, depth)
        work = _BUILD_WORK.get(key)             # the bitmaps: reused from frame to frame (stream-ordered use)
        if work is None:
            work = _BUILD_WORK[key] = torch.empty((lib.svoxb_build_dense_work_bytes(depth),), dtype=torch.uint8, device=dev)
        status = torch.zeros((2,), dtype=torch.int64, device=dev)

        def run(cap):
            child = torch.empty((cap, 2, 2, 2), dtype=torch.int32, device=dev)
            data = torch.empty((cap, 2, 2, 2, 1), dtype=torch.int32, device=dev)
            parent_depth = torch.empty((cap, 2), dtype=torch.int32, device=dev)
            _check(lib.svoxb_build_dense(_ptr(points), P, depth, _ptr(offset), _ptr(scaling), _ptr(work), cap, _ptr(child),
                                         _ptr(data), _ptr(parent_depth), _ptr(status), _stream()))
            return child, data, parent_depth
        if capacity is not None:
            return (*run(int(capacity)), status)
        # exact size: a first pass with an upper bound ((depth-1) internal nodes per point + the root, and no more than
        # a full tree), one read-back, then trim (views of the same storage)
        bound = min(1 + P * max(depth - 1, 0), sum(8 ** l for l in range(depth)))
        child, data, parent_depth = run(max(1, min(bound, 1 << 22)))
        need, over = (int(v) for v in status.tolist())
        if over:
            child, data, parent_depth = run(need)
        child, data, parent_depth = child[:need], data[:need], parent_depth[:need]
        if child.untyped_storage().nbytes() > 2 * child.numel() * 4:      # do not pin the upper-bound allocation
            child, data, parent_depth = child.clone(), data.clone(), parent_depth.clone()
        return child, data, parent_depth, None


def _unsupported(name, why):
    def f(*a, **k):
        raise RuntimeError(f"svox_t_b200.csrc.{name} is not implemented: {why}")
    f.__name__ = name
    return f


def query_vertical_backward(tree, indices, grad_output):
    """grad_data[M, K]: row scatter-add of ``grad_output[Q, K]`` into the rows the points fall in
    (svox_kernel.cu:380-403; the reference's kernel faults, Appendix B1 -- this is what its source states)."""
    lib = load_library()
    _check_input(indices, "indices", torch.float32)
    _check_input(grad_output, "grad_output", torch.float32)
    if indices.dim() != 2 or indices.shape[1] != 3 or grad_output.dim() != 2 or grad_output.shape[0] != indices.shape[0]:
        raise RuntimeError("indices must be [Q, 3] and grad_output [Q, K]")
    ct = tree._c()
    dev = indices.device
    with torch.cuda.device(dev):
        grad = torch.zeros((tree.features.shape[0], grad_output.shape[1]), dtype=torch.float32, device=dev)
        _check(lib.svoxb_query_bwd(ctypes.byref(ct), _ptr(indices), indices.shape[0], _ptr(grad_output),
                                   grad_output.shape[1], _ptr(grad), _stream()))
    return grad


def assign_vertical(tree, indices, values):
    """features[row(p_q), :K] = values[q], in place (svox_kernel.cu:326-339). Deterministic: the largest q wins a
    shared leaf."""
    lib = load_library()
    _check_input(indices, "indices", torch.float32)
    _check_input(values, "values", torch.float32)
    if indices.dim() != 2 or indices.shape[1] != 3 or values.dim() != 2 or values.shape[0] != indices.shape[0]:
        raise RuntimeError("indices must be [Q, 3] and values [Q, K]")
    ct = tree._c()
    with torch.cuda.device(indices.device):
        _check(lib.svoxb_assign(ctypes.byref(ct), _ptr(tree.features), _ptr(indices), indices.shape[0], _ptr(values),
                                values.shape[1], _stream()))
    # the kernel wrote through the raw pointer: bump the tensor version so that activated tables / hit marks derived
    # from the old rows are rebuilt
    torch.autograd.graph.increment_version(tree.features)


def calc_corners(tree, indexer):
    """[Q, 3] lower corners (tree coordinates) of the cells ``indexer[Q, 4] = [node, i, j, k]``
    (svox_kernel.cu:436-457; the reference dispatches on its int32 ``data`` tensor and raises, Appendix B4)."""
    lib = load_library()
    _check_input(indexer, "indexer", torch.int64)
    if indexer.dim() != 2 or indexer.shape[1] != 4:
        raise RuntimeError("indexer must be [Q, 4]")
    _check_input(tree.parent_depth, "parent_depth", torch.int32)
    dev = indexer.device
    with torch.cuda.device(dev):
        out = torch.empty((indexer.shape[0], 3), dtype=torch.float32, device=dev)
        _check(lib.svoxb_calc_corners(_ptr(tree.parent_depth), tree.child.shape[1], int(tree.n_internal),
                                      _ptr(indexer), indexer.shape[0], _ptr(out), _stream()))
    return out


def grid_weight_render(data, cam, opt, offset, scaling):
    """[grid_weight, grid_hit], both shaped like the dense sigma grid ``data[r, r, r]``: per cell the largest compositing
    weight any pixel ray of ``cam`` leaves there and the number of hits (rt_kernel.cu:1454-1478)."""
    lib = load_library()
    for t, n in ((data, "data"), (offset, "offset"), (scaling, "scaling")):
        _check_input(t, n, torch.float32)
    if data.dim() != 3 or data.shape[0] != data.shape[1] or data.shape[0] != data.shape[2]:
        raise RuntimeError("data must be a cubic [r, r, r] grid")
    cam.check()
    cc = cam._c()
    dev = data.device
    with torch.cuda.device(dev):
        gw, gh = torch.zeros_like(data), torch.zeros_like(data)
        _check(lib.svoxb_grid_weight_render(_ptr(data), data.shape[0], ctypes.byref(cc), ctypes.byref(opt._c()),
                                            _ptr(offset), _ptr(scaling), _ptr(gw), _ptr(gh), _stream()))
    return [gw, gh]


# The one reference entry point that stays out: present so that `hasattr(_C, name)` behaves, but
# it raises instead of silently doing something else.
quantize_median_cut = _unsupported("quantize_median_cut", "CPU-only PlenOctree leftover; out of scope")
