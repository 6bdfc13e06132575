"""ctypes front-end of the CPU oracle (oracle/svox_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs. The product package (svox_t_b200) must never import this module.

All entry points take/return numpy arrays. ``dtype`` selects the arithmetic: ``np.float32`` restates the
reference's fp32 CUDA instantiation, ``np.float64`` is the all-double mathematical truth.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_SO = os.path.join(_BUILD, "libsvox_oracle.so")
_SRC = [os.path.join(_HERE, "svox_oracle.c"), os.path.join(_HERE, "svox_oracle_impl.inc")]
_lock = threading.Lock()
_lib = None

SENTINEL = 1410065408  # int(1e10) wrapped to int32 (svox_t/svox.py:124)


def _cpu_has_fma():
    try:
        with open("/proc/cpuinfo") as f:
            return " fma " in f.read().replace("\n", " ")
    except OSError:
        return False


def build(force=False):
    """Compile the C restatement with gcc (about one second). Idempotent."""
    os.makedirs(_BUILD, exist_ok=True)
    flag_file = os.path.join(_BUILD, "flags.txt")
    flags = ["-O2", "-fopenmp", "-ffp-contract=off"] + (["-mfma"] if _cpu_has_fma() else [])
    want = " ".join(flags)
    stale = force or not os.path.exists(_SO) or not os.path.exists(flag_file)
    if not stale:
        stale = open(flag_file).read().strip() != want or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in _SRC)
    if stale:
        tmp = _SO + ".tmp.%d" % os.getpid()
        subprocess.check_call(["gcc", *flags, "-shared", "-fPIC", _SRC[0], "-o", tmp, "-lm"])
        os.replace(tmp, _SO)
        with open(flag_file, "w") as f:
            f.write(want)
    return _SO


def lib():
    global _lib
    with _lock:
        if _lib is None:
            _lib = ctypes.CDLL(build())
            _lib.orc_leafset.restype = ctypes.c_int64
    return _lib


def set_threads(n):
    """Use ``n`` OpenMP threads in the oracle's loops (returns the count in effect); see orc_set_threads."""
    return int(lib().orc_set_threads(ctypes.c_int(int(n))))


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError(dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


class Tree:
    """Plain container of the reference tensor format (svox_t/svox.py:121-139) as numpy arrays."""

    def __init__(self, child, data, offset=(0.0, 0.0, 0.0), scaling=(1.0, 1.0, 1.0)):
        self.child = _c(child, np.int32)
        self.data = _c(data, np.int32).reshape(self.child.shape)
        self.N = int(self.child.shape[1])
        self.offset = np.asarray(offset, dtype=np.float64)
        self.scaling = np.asarray(scaling, dtype=np.float64)

    def args(self, features, dtype):
        f = _c(features, dtype)
        M, D = f.shape
        off, sc = _c(self.offset, dtype), _c(self.scaling, dtype)
        keep = (f, off, sc)
        return keep, (_p(self.child), _p(self.data), ctypes.c_int(self.N), _p(f), ctypes.c_int64(M),
                      ctypes.c_int(D), _p(off), _p(sc))


def query(tree, features, pts, dtype=np.float32):
    """-> values[Q,D], node_ids[Q] i64, data_ids[Q] i64, valid[Q] bool (rows of empty leaves are zero)."""
    sfx, _ = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    pts = _c(pts, dtype)
    Q, D = pts.shape[0], keep[0].shape[1]
    values = np.zeros((Q, D), dtype=dtype)
    node_ids = np.zeros(Q, dtype=np.int64)
    data_ids = np.zeros(Q, dtype=np.int64)
    valid = np.zeros(Q, dtype=np.uint8)
    getattr(lib(), "orc_query" + sfx)(*targs, _p(pts), ctypes.c_int64(Q), _p(values), _p(node_ids),
                                       _p(data_ids), _p(valid))
    return values, node_ids, data_ids, valid.astype(bool)


def leafset(node_ids, N):
    node_ids = _c(node_ids, np.int64)
    out = np.zeros((max(len(node_ids), 1), 4), dtype=np.int64)
    n = lib().orc_leafset(_p(node_ids), ctypes.c_int64(len(node_ids)), ctypes.c_int(N), _p(out))
    return out[:n].copy()


def construct_tree(tree, pts, dtype=np.float32):
    """In place: tree.data[leaf(p_i)] = i (highest i wins on a shared leaf)."""
    sfx, _ = _sfx(dtype)
    pts = _c(pts, dtype)
    off, sc = _c(tree.offset, dtype), _c(tree.scaling, dtype)
    getattr(lib(), "orc_construct_tree" + sfx)(_p(tree.child), _p(tree.data), ctypes.c_int(tree.N),
                                                _p(off), _p(sc), _p(pts), ctypes.c_int64(len(pts)))


def render_rays(tree, features, origins, dirs, step_size=1e-3, background_brightness=1.0,
                sigma_thresh=0.0, stop_thresh=0.0, dtype=np.float32, want_depth=True, want_counters=False):
    """-> out[Q,D] (D-1 composited sigmoid features + opacity), depth[Q], [counters dict S, LV, V, H]."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d = _c(origins, dtype), _c(dirs, dtype)
    Q, D = o.shape[0], keep[0].shape[1]
    out = np.zeros((Q, D), dtype=dtype)
    depth = np.zeros(Q, dtype=dtype) if want_depth else None
    cnt = np.zeros(4, dtype=np.int64)
    getattr(lib(), "orc_render_rays" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(Q), cr(step_size),
                                             cr(background_brightness), cr(sigma_thresh), cr(stop_thresh),
                                             _p(out), _p(depth), _p(cnt))
    ret = [out, depth]
    if want_counters:
        ret.append(dict(S=int(cnt[0]), LV=int(cnt[1]), V=int(cnt[2]), H=int(cnt[3]), Q=int(Q)))
    return tuple(ret)


def render_rays_backward(tree, features, origins, dirs, grad_out, step_size=1e-3,
                         background_brightness=1.0, dtype=np.float32):
    """-> grad[M,D] following the reference's two-pass backward (hit predicate sigma > 0, no early stop)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d, g = _c(origins, dtype), _c(dirs, dtype), _c(grad_out, dtype)
    grad = np.zeros_like(keep[0])
    getattr(lib(), "orc_render_rays_backward" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]),
                                                      cr(step_size), cr(background_brightness), _p(g), _p(grad))
    return grad


def opacity_render(tree, features, origins, dirs, step_size=1e-3, sigma_thresh=0.0, stop_thresh=0.0,
                   dtype=np.float32):
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d = _c(origins, dtype), _c(dirs, dtype)
    out = np.zeros(o.shape[0], dtype=dtype)
    getattr(lib(), "orc_opacity_render" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]), cr(step_size),
                                                cr(sigma_thresh), cr(stop_thresh), _p(out))
    return out


def opacity_render_backward(tree, features, origins, dirs, grad_out, step_size=1e-3, dtype=np.float32):
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d, g = _c(origins, dtype), _c(dirs, dtype), _c(grad_out, dtype).reshape(-1)
    grad = np.zeros_like(keep[0])
    getattr(lib(), "orc_opacity_render_backward" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]),
                                                         cr(step_size), _p(g), _p(grad))
    return grad


def motion_render(tree, features, origins, dirs, extra, step_size=1e-3, sigma_thresh=0.0, dtype=np.float32):
    """-> out[Q,J], depth[Q], hit_point[Q,3], data_idx[Q] i64 (zeros where nothing is hit)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d, e = _c(origins, dtype), _c(dirs, dtype), _c(extra, dtype)
    Q, J = o.shape[0], e.shape[0]
    out = np.zeros((Q, J), dtype=dtype)
    depth = np.zeros(Q, dtype=dtype)
    hit = np.zeros((Q, 3), dtype=dtype)
    didx = np.zeros(Q, dtype=np.int64)
    getattr(lib(), "orc_motion_render" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(Q), cr(step_size),
                                               cr(sigma_thresh), _p(e), ctypes.c_int(J), _p(out), _p(depth), _p(hit),
                                               _p(didx))
    return out, depth, hit, didx


def camera_rays(c2w, fx, fy, width, height, dtype=np.float32):
    """-> origins[H*W,3], dirs[H*W,3] in row-major pixel order (iy*W + ix), world space."""
    sfx, cr = _sfx(dtype)
    c = np.zeros((3, 4), dtype=dtype)
    c[:] = np.asarray(c2w, dtype=dtype)[:3, :4]
    n = width * height
    o = np.zeros((n, 3), dtype=dtype)
    d = np.zeros((n, 3), dtype=dtype)
    getattr(lib(), "orc_camera_rays" + sfx)(_p(c), cr(fx), cr(fy), ctypes.c_int(width), ctypes.c_int(height),
                                             _p(o), _p(d))
    return o, d


def warp_vertices(T, coords, weights, joint_index, dtype=np.float32):
    sfx, _ = _sfx(dtype)
    T, coords, weights = _c(T, dtype), _c(coords, dtype), _c(weights, dtype)
    ji = _c(joint_index, np.int32)
    P, B = weights.shape
    co = np.zeros((P, 3), dtype=dtype)
    mo = np.zeros((P, 4, 4), dtype=dtype)
    getattr(lib(), "orc_warp_vertices" + sfx)(_p(T), _p(coords), _p(weights), _p(ji), ctypes.c_int64(P),
                                               ctypes.c_int(B), _p(co), _p(mo))
    return co, mo


def p2v(points, feat, corner, size, n_voxels, kernel_radius, conv_radius, dtype=np.float32):
    sfx, cr = _sfx(dtype)
    points, feat = _c(points, dtype), _c(feat, dtype)
    corner, size = _c(corner, dtype), _c(size, dtype)
    vox = np.zeros((n_voxels, n_voxels, n_voxels, 1), dtype=dtype)
    getattr(lib(), "orc_p2v" + sfx)(_p(points), _p(feat), ctypes.c_int64(points.shape[0]),
                                     ctypes.c_int(feat.shape[1]), _p(corner), _p(size), ctypes.c_int(n_voxels),
                                     cr(kernel_radius), cr(conv_radius), _p(vox))
    return vox


def warp_vertices_backward(T, coords, weights, joint_index, g_coords, g_mats, dtype=np.float32):
    """-> grad_coords[P,3], grad_T[J,4,4], grad_w[P,B]."""
    sfx, _ = _sfx(dtype)
    T, coords, weights = _c(T, dtype), _c(coords, dtype), _c(weights, dtype)
    g_coords, g_mats = _c(g_coords, dtype), _c(g_mats, dtype)
    ji = _c(joint_index, np.int32)
    P, B = weights.shape
    gT, gx, gw = np.zeros_like(T), np.zeros((P, 3), dtype=dtype), np.zeros((P, B), dtype=dtype)
    getattr(lib(), "orc_warp_vertices_backward" + sfx)(_p(T), _p(coords), _p(weights), _p(ji), _p(g_coords), _p(g_mats),
                                                        ctypes.c_int64(P), ctypes.c_int(B), _p(gT), _p(gx), _p(gw))
    return gx, gT, gw


def p2v_backward(g_vox, points, feat, corner, size, n_voxels, kernel_radius, conv_radius, dtype=np.float32):
    """-> grad_points[P,3], grad_feat[P,F] (value in channel 0, as the reference writes it)."""
    sfx, cr = _sfx(dtype)
    g_vox, points, feat = _c(g_vox, dtype), _c(points, dtype), _c(feat, dtype)
    corner, size = _c(corner, dtype), _c(size, dtype)
    gp, gf = np.zeros_like(points), np.zeros_like(feat)
    getattr(lib(), "orc_p2v_backward" + sfx)(_p(g_vox), _p(points), _p(feat), ctypes.c_int64(points.shape[0]),
                                              ctypes.c_int(feat.shape[1]), _p(corner), _p(size), ctypes.c_int(n_voxels),
                                              cr(kernel_radius), cr(conv_radius), _p(gp), _p(gf))
    return gp, gf


# ---- view-dependent formats, NDC cameras, motion-feature render ---------------------------------
FORMAT_RGBA, FORMAT_SH, FORMAT_SG, FORMAT_ASG = 0, 1, 2, 3


def _fmt_args(fmt, basis_dim, min_comp, max_comp, extra, tm, dtype):
    max_comp = max_comp if max_comp >= 0 else max_comp + basis_dim
    e = None if extra is None else _c(extra, dtype)
    t = None if tm is None else _c(tm, dtype)
    cols = 0 if e is None else int(e.shape[1])
    keep = (e, t)
    return keep, (ctypes.c_int(fmt), ctypes.c_int(basis_dim), ctypes.c_int(min_comp), ctypes.c_int(max_comp),
                  _p(e), ctypes.c_int(cols), _p(t))


def render_rays_fmt(tree, features, origins, dirs, vdirs, fmt, basis_dim, extra=None, tm=None, min_comp=0,
                    max_comp=-1, step_size=1e-3, background_brightness=1.0, sigma_thresh=0.0, stop_thresh=0.0,
                    dtype=np.float32):
    """SH / SG / ASG render -> out[Q, C+1], C = (D-1)//basis_dim (rt_kernel.cu:293-301). ``tm`` [M,4,4] rotates the
    view direction per hit row (rt_kernel.cu:283-291)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    keep2, fargs = _fmt_args(fmt, basis_dim, min_comp, max_comp, extra, tm, dtype)
    o, d, v = _c(origins, dtype), _c(dirs, dtype), _c(vdirs, dtype)
    Q, D = o.shape[0], keep[0].shape[1]
    out = np.zeros((Q, (D - 1) // basis_dim + 1), dtype=dtype)
    getattr(lib(), "orc_render_rays_fmt" + sfx)(*targs, _p(o), _p(d), _p(v), ctypes.c_int64(Q), cr(step_size),
                                                 cr(background_brightness), cr(sigma_thresh), cr(stop_thresh),
                                                 *fargs, _p(out))
    return out


def render_rays_fmt_backward(tree, features, origins, dirs, vdirs, grad_out, fmt, basis_dim, extra=None, tm=None,
                             min_comp=0, max_comp=-1, step_size=1e-3, background_brightness=1.0, stale_basis=False,
                             dtype=np.float32):
    """-> grad[M,D]. stale_basis=True reproduces the reference's second pass with per-row rotations (see the C file)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    keep2, fargs = _fmt_args(fmt, basis_dim, min_comp, max_comp, extra, tm, dtype)
    o, d, v, g = _c(origins, dtype), _c(dirs, dtype), _c(vdirs, dtype), _c(grad_out, dtype)
    grad = np.zeros_like(keep[0])
    getattr(lib(), "orc_render_rays_fmt_backward" + sfx)(*targs, _p(o), _p(d), _p(v), ctypes.c_int64(o.shape[0]),
                                                          cr(step_size), cr(background_brightness), *fargs, _p(g),
                                                          ctypes.c_int(1 if stale_basis else 0), _p(grad))
    return grad


def camera_rays_ndc(c2w, fx, fy, width, height, ndc_width=-1, ndc_height=-1, ndc_focal=0.0, dtype=np.float32):
    """-> origins, dirs (NDC when ndc_width >= 0), vdirs (world) of the image kernels (rt_kernel.cu:1193-1206)."""
    sfx, cr = _sfx(dtype)
    c = np.zeros((3, 4), dtype=dtype)
    c[:] = np.asarray(c2w, dtype=dtype)[:3, :4]
    n = width * height
    o, d, v = (np.zeros((n, 3), dtype=dtype) for _ in range(3))
    getattr(lib(), "orc_camera_rays_ndc" + sfx)(_p(c), cr(fx), cr(fy), ctypes.c_int(width), ctypes.c_int(height),
                                                 ctypes.c_int(ndc_width), ctypes.c_int(ndc_height), cr(ndc_focal),
                                                 _p(o), _p(d), _p(v))
    return o, d, v


def motion_feature_render(tree, features, origins, dirs, joint_features, skinning_weights, joint_index,
                          step_size=1e-3, background_brightness=1.0, sigma_thresh=0.0, stop_thresh=0.0,
                          dtype=np.float32):
    """-> out[Q,F] (rt_kernel.cu:885-979)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d = _c(origins, dtype), _c(dirs, dtype)
    jf, sw, ji = _c(joint_features, dtype), _c(skinning_weights, dtype), _c(joint_index, np.int32)
    F, B = jf.shape[1], sw.shape[1]
    assert F <= 64
    out = np.zeros((o.shape[0], F), dtype=dtype)
    getattr(lib(), "orc_motion_feature_render" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]), cr(step_size),
                                                       cr(background_brightness), cr(sigma_thresh), cr(stop_thresh),
                                                       _p(jf), _p(sw), _p(ji), ctypes.c_int(F), ctypes.c_int(B), _p(out))
    return out


def motion_feature_render_backward(tree, features, origins, dirs, joint_features, skinning_weights, joint_index,
                                   grad_out, step_size=1e-3, dtype=np.float32):
    """-> grad_joint_features[J,F]: the gradient the reference meant to compute (Appendix B3)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d, g = _c(origins, dtype), _c(dirs, dtype), _c(grad_out, dtype)
    jf, sw, ji = _c(joint_features, dtype), _c(skinning_weights, dtype), _c(joint_index, np.int32)
    F, B = jf.shape[1], sw.shape[1]
    grad = np.zeros_like(jf)
    getattr(lib(), "orc_motion_feature_render_backward" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]),
                                                                cr(step_size), _p(jf), _p(sw), _p(ji), ctypes.c_int(F),
                                                                ctypes.c_int(B), _p(g), _p(grad))
    return grad


def accumulate_weights(tree, features, origins, dirs, step_size=1e-3, sigma_thresh=0.0, stop_thresh=0.0,
                       dtype=np.float32):
    """-> weight_accum with the shape of child: per-leaf sum of the compositing weights (rt_kernel.cu:308-310)."""
    sfx, cr = _sfx(dtype)
    keep, targs = tree.args(features, dtype)
    o, d = _c(origins, dtype), _c(dirs, dtype)
    wa = np.zeros(tree.child.shape, dtype=dtype)
    getattr(lib(), "orc_accumulate_weights" + sfx)(*targs, _p(o), _p(d), ctypes.c_int64(o.shape[0]), cr(step_size),
                                                    cr(sigma_thresh), cr(stop_thresh), _p(wa))
    return wa


def query_backward(tree, M, pts, grad_out, dtype=np.float32):
    """query_vertical_backward as the reference's source states it -> grad_data[M,K]."""
    sfx, _ = _sfx(dtype)
    pts, g = _c(pts, dtype), _c(grad_out, dtype)
    off, sc = _c(tree.offset, dtype), _c(tree.scaling, dtype)
    out = np.zeros((M, g.shape[1]), dtype=dtype)
    getattr(lib(), "orc_query_backward" + sfx)(_p(tree.child), _p(tree.data), ctypes.c_int(tree.N), ctypes.c_int64(M),
                                                _p(off), _p(sc), _p(pts), ctypes.c_int64(len(pts)), _p(g),
                                                ctypes.c_int(g.shape[1]), _p(out))
    return out


def assign(tree, features, pts, values, dtype=np.float32):
    """assign_vertical (highest point index wins a shared leaf) -> the updated copy of features."""
    sfx, _ = _sfx(dtype)
    f = np.array(features, dtype=dtype, order="C", copy=True)
    pts, v = _c(pts, dtype), _c(values, dtype)
    off, sc = _c(tree.offset, dtype), _c(tree.scaling, dtype)
    getattr(lib(), "orc_assign" + sfx)(_p(tree.child), _p(tree.data), ctypes.c_int(tree.N), _p(f),
                                        ctypes.c_int64(f.shape[0]), ctypes.c_int(f.shape[1]), _p(off), _p(sc),
                                        _p(pts), ctypes.c_int64(len(pts)), _p(v), ctypes.c_int(v.shape[1]))
    return f


def calc_corners(parent_depth, N, indexer, dtype=np.float32):
    sfx, _ = _sfx(dtype)
    pd, ix = _c(parent_depth, np.int32), _c(indexer, np.int64)
    out = np.zeros((len(ix), 3), dtype=dtype)
    getattr(lib(), "orc_calc_corners" + sfx)(_p(pd), ctypes.c_int(N), _p(ix), ctypes.c_int64(len(ix)), _p(out))
    return out


def grid_weight_render(grid, c2w, fx, fy, width, height, offset=(0.0, 0.0, 0.0), scaling=(1.0, 1.0, 1.0),
                       step_size=1e-3, sigma_thresh=0.0, ndc_width=-1, ndc_height=-1, ndc_focal=0.0,
                       dtype=np.float32):
    """-> grid_weight[r,r,r], grid_hit[r,r,r]."""
    sfx, ct = _sfx(dtype)
    g = _c(grid, dtype)
    c = np.zeros((4, 4), dtype=dtype)
    c2w = np.asarray(c2w, dtype=dtype)
    c[:c2w.shape[0]] = c2w
    off, sc = _c(np.asarray(offset), dtype), _c(np.asarray(scaling), dtype)
    gw, gh = np.zeros_like(g), np.zeros_like(g)
    getattr(lib(), "orc_grid_weight_render" + sfx)(
        _p(g), ctypes.c_int(g.shape[0]), _p(c), ct(fx), ct(fy), ctypes.c_int(width), ctypes.c_int(height),
        ctypes.c_int(ndc_width), ctypes.c_int(ndc_height), ct(ndc_focal), _p(off), _p(sc), ct(step_size),
        ct(sigma_thresh), _p(gw), _p(gh))
    return gw, gh
