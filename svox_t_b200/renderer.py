"""VolumeRenderer -- differentiable feature-level volume rendering of an N3Tree.

API of the reference's ``svox_t.renderer`` (svox_t/renderer.py:162-439): ``forward(features, rays, ...)``,
``render_persp(features, c2w, width, height, fx, fy)``, ``render_depth(features, rays)``; output rows are
``D-1`` composited sigmoid features followed by opacity (``data_format`` RGBA, the hot path) or
``(D-1)/basis_dim`` view-dependent channels followed by opacity (SH / SG / ASG trees). Gradients flow into
``features`` through a custom autograd Function backed by the single-re-march backward kernel.

Extensions: ``forward_with_depth`` / ``render_persp_with_depth`` return the first-hit depth from the same march
(the reference launches a second kernel, renderer.py:377-382).
"""
from collections import namedtuple
from warnings import warn

import torch
from torch import autograd, nn

from . import csrc as _C
from .helpers import DataFormat

NDCConfig = namedtuple("NDCConfig", ["width", "height", "focal"])
Rays = namedtuple("Rays", ["origins", "dirs", "viewdirs"])


def _rays_spec_from_rays(rays):
    spec = _C.RaysSpec()
    spec.origins = rays.origins
    spec.dirs = rays.dirs
    spec.vdirs = rays.viewdirs
    return spec


def _make_camera_spec(c2w, width, height, fx, fy, rows=None):
    spec = _C.CameraSpec()
    spec.c2w = c2w
    spec.width = width
    spec.height = height
    spec.fx = fx
    spec.fy = fy
    if rows is not None:
        spec.row_begin, spec.row_end = int(rows[0]), int(rows[1])
    return spec


class _VolumeRenderFunction(autograd.Function):
    """renderer.py:60-77; additionally keeps the forward output, which the one-pass backward consumes."""

    @staticmethod
    def forward(ctx, data, tree, rays, opt, want_depth):
        out, depth = _C._render_fwd(tree, rays, opt, want_depth)
        ctx.tree, ctx.rays, ctx.opt = tree, rays, opt
        # save_for_backward, not a ctx attribute: an output kept as an attribute forms a reference cycle
        # (out -> grad_fn -> ctx -> out) that only the cyclic GC frees -- hundreds of MB per step linger, and the
        # caching allocator answers with sporadic 50 ms cudaMallocs.
        ctx.save_for_backward(out)
        if want_depth:
            ctx.mark_non_differentiable(depth)
            return out, depth
        return out

    @staticmethod
    def backward(ctx, grad_out, *_unused):
        if ctx.needs_input_grad[0]:
            return (_C.volume_render_backward(ctx.tree, ctx.rays, ctx.opt, grad_out.contiguous(),
                                              saved_out=ctx.saved_tensors[0]), None, None, None, None)
        return None, None, None, None, None


class _VolumeRenderImageFunction(autograd.Function):
    """renderer.py:79-94."""

    @staticmethod
    def forward(ctx, data, tree, cam, opt, want_depth):
        out, depth = _C._render_image_fwd(tree, cam, opt, want_depth)
        ctx.tree, ctx.cam, ctx.opt = tree, cam, opt
        ctx.save_for_backward(out)
        if want_depth:
            ctx.mark_non_differentiable(depth)
            return out, depth
        return out

    @staticmethod
    def backward(ctx, grad_out, *_unused):
        if ctx.needs_input_grad[0]:
            return (_C.volume_render_image_backward(ctx.tree, ctx.cam, ctx.opt, grad_out.contiguous(),
                                                    saved_out=ctx.saved_tensors[0]), None, None, None, None)
        return None, None, None, None, None


class _MotionFeatureRenderFunction(autograd.Function):
    """renderer.py:96-116; differentiable w.r.t. ``joint_features`` with the gradient the reference meant to compute
    (its kernel is broken)."""

    @staticmethod
    def forward(ctx, data, tree, rays, opt):
        ctx.tree, ctx.rays, ctx.opt = tree, rays, opt
        return _C.motion_feature_render(tree, rays, opt)

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.needs_input_grad[0]:
            return (_C.motion_feature_render_backward(ctx.tree, ctx.rays, ctx.opt, grad_out.contiguous()),
                    None, None, None)
        return None, None, None, None


class _OpacityRenderFunction(autograd.Function):
    """renderer.py:118-138, with the backward the reference meant to run (Appendix B2)."""

    @staticmethod
    def forward(ctx, data, tree, rays, opt):
        ctx.tree, ctx.rays, ctx.opt = tree, rays, opt
        out = _C.opacity_render(tree, rays, opt)
        ctx.save_for_backward(out)          # T_end = 1 - out: the backward marches once instead of twice
        return out

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.needs_input_grad[0]:
            return (_C.opacity_render_backward(ctx.tree, ctx.rays, ctx.opt, grad_out.contiguous(),
                                               saved_out=ctx.saved_tensors[0]), None, None, None)
        return None, None, None, None


class VolumeRenderer(nn.Module):
    """Volume renderer bound to an N3Tree (renderer.py:162-205)."""

    def __init__(self, tree, step_size: float = 1e-3, background_brightness: float = 1.0, ndc: NDCConfig = None,
                 min_comp=0, max_comp=-1):
        super().__init__()
        self.tree = tree
        self.step_size = step_size
        self.background_brightness = background_brightness
        self.ndc_config = ndc
        self.min_comp = min_comp
        self.max_comp = max_comp
        if isinstance(tree.data_format, DataFormat):
            self.data_format = tree.data_format
        else:
            warn("N3Tree without data_format, assuming the feature-level RGBA format")
            self.data_format = DataFormat("RGBA")
        if self.max_comp < 0:
            self.max_comp += self.data_format.basis_dim
        self.tree._weight_accum = None
        # multi-GPU training (svox_t_b200 extension): set to a dist.LeafGradExchange and backward() leaves the leaf
        # gradients of forward() summed over the GPUs, in the exchange's table (valid until the next backward)
        self.leaf_grad_exchange = None

    def _render_spec(self, features, n_rays, **kw):
        """TreeSpec for a march over ``n_rays`` rays. When the batch is large enough for every leaf row to be visited
        many times, attach the pre-activated table (sigmoid once per row instead of once per visit); small batches
        keep the in-kernel sigmoid, for which one pass over the whole table would cost more than it saves."""
        ts = self.tree._spec(features, **kw)
        ts._grad_exchange = getattr(self, "leaf_grad_exchange", None)
        M, D = features.shape
        if n_rays * 32 >= M and features.is_cuda and features.dtype == torch.float32:
            if self.data_format.format == DataFormat.RGBA and 2 <= D <= 128:
                # one pass over the rows: activated table + hit marks (+ zero-fill of the exchange's gradient table when
                # this forward will be back-propagated)
                xchg = ts._grad_exchange if (features.requires_grad and torch.is_grad_enabled()) else None
                ts._act = self.tree.activated(features.detach(), accel=ts._accel, grad_exchange=xchg)
            elif ts._accel is not None:     # every format keeps sigma in the last channel: dead rows are never fetched
                ts._accel.mark_hits(features.detach())
        return ts

    def _sigma_spec(self, features, n_rays):
        """TreeSpec for the marches that only read sigma (depth, opacity, motion): for batches large enough to pay for
        two small passes, attach the compact sigma array and refresh the hit marks (dead rows are never fetched)."""
        ts = self.tree._spec(features)
        if n_rays * 32 >= features.shape[0] and features.is_cuda and features.dtype == torch.float32:
            ts._sigma = self.tree.sigma_table(features.detach())
            if ts._accel is not None:
                ts._accel.mark_hits(features.detach())
        return ts

    def _require_cuda(self, cuda):
        if not cuda or not self.tree.data.is_cuda:
            # the reference asserts False here (renderer.py:225,335): its PyTorch path is dead code
            raise RuntimeError("svox_t_b200 renders on CUDA only: there is no CPU / PyTorch fallback path")

    def forward(self, features, rays: Rays, transformation_matrices=None, cuda=True, fast=False):
        """Render a ray batch -> (B, D): D-1 features + opacity (RGBA), or (B, (D-1)/basis_dim + 1) for SH / SG / ASG
        trees. Differentiable w.r.t. ``features``. ``transformation_matrices`` [M,4,4] rotates the view direction per
        hit row before the basis is evaluated; as in the reference it has no effect on the RGBA format."""
        self._require_cuda(cuda)
        return _VolumeRenderFunction.apply(
            features, self._render_spec(features, rays.origins.shape[0], transformation_matrices=transformation_matrices),
            _rays_spec_from_rays(rays), self._get_options(fast), False)

    def forward_with_depth(self, features, rays: Rays, fast=False):
        """(out (B, D), depth (B, 1)) from one march."""
        self._require_cuda(True)
        return _VolumeRenderFunction.apply(features, self._render_spec(features, rays.origins.shape[0]),
                                           _rays_spec_from_rays(rays), self._get_options(fast), True)

    def render_persp(self, features, c2w, width=800, height=800, fx=1111.111, fy=None, cuda=True, fast=False,
                     rows=None):
        """Perspective image -> (height, width, D). Differentiable (renderer.py:310-366).
        ``rows=(y0, y1)`` (svox_t_b200 extension) renders only that band of image rows -> (y1 - y0, width, D): one
        frame split over several GPUs (svox_t_b200.dist.render_image_bands)."""
        self._require_cuda(cuda)
        fy = fx if fy is None else fy
        n_rows = height if rows is None else rows[1] - rows[0]
        return _VolumeRenderImageFunction.apply(features, self._render_spec(features, width * n_rows),
                                                _make_camera_spec(c2w, width, height, fx, fy, rows),
                                                self._get_options(fast), False)

    def render_persp_with_depth(self, features, c2w, width=800, height=800, fx=1111.111, fy=None, fast=False):
        """(image (H, W, D), depth (H, W, 1)) from one march."""
        self._require_cuda(True)
        fy = fx if fy is None else fy
        return _VolumeRenderImageFunction.apply(features, self._render_spec(features, width * height),
                                                _make_camera_spec(c2w, width, height, fx, fy),
                                                self._get_options(fast), True)

    def render_depth(self, features, rays: Rays, cuda=True, fast=False):
        """First-hit depth (B, 1), not differentiable (renderer.py:377-382)."""
        self._require_cuda(cuda)
        return _C.render_depth(self._sigma_spec(features, rays.origins.shape[0]), _rays_spec_from_rays(rays),
                               self._get_options(fast))

    def motion_render(self, features, rays: Rays, cuda=True, fast=False):
        """First-hit joint distances, depth, hit point and data index (renderer.py:367-375)."""
        assert self.tree.extra_data is not None, "Need extra data to store skeleton postion."
        self._require_cuda(cuda)
        return tuple(_C.motion_render(self._sigma_spec(features, rays.origins.shape[0]), _rays_spec_from_rays(rays),
                                      self._get_options(fast)))

    def motion_feature_render(self, features, joint_features, skinning_weights, joint_index, rays: Rays, cuda=True,
                              fast=False):
        """Composited per-joint features (B, F): every hit blends ``joint_features`` [J,F] with the hit row's
        ``skinning_weights`` / ``joint_index`` [M,B]. Differentiable w.r.t. ``joint_features`` (renderer.py:384-396)."""
        self._require_cuda(cuda)
        ts = self.tree._spec(features, joint_features, skinning_weights, joint_index)
        if ts._accel is not None and rays.origins.shape[0] * 32 >= features.shape[0]:
            ts._accel.mark_hits(features.detach())       # rows with sigma <= 0 never become candidates of the march
        return _MotionFeatureRenderFunction.apply(joint_features, ts, _rays_spec_from_rays(rays), self._get_options(fast))

    def opacity_render(self, features, rays: Rays, cuda=True, fast=False):
        """Opacity only (B, 1); differentiable w.r.t. the sigma channel of ``features`` (renderer.py:397-406)."""
        self._require_cuda(cuda)
        return _OpacityRenderFunction.apply(features, self._sigma_spec(features, rays.origins.shape[0]),
                                            _rays_spec_from_rays(rays), self._get_options(fast))

    def _get_options(self, fast=False):
        """RenderOptions for the kernels (renderer.py:408-439)."""
        opts = _C.RenderOptions()
        opts.step_size = self.step_size
        opts.background_brightness = self.background_brightness
        opts.format = self.data_format.format
        opts.basis_dim = self.data_format.basis_dim
        opts.min_comp = self.min_comp
        opts.max_comp = self.max_comp
        if self.ndc_config is not None:
            opts.ndc_width = self.ndc_config.width
            opts.ndc_height = self.ndc_config.height
            opts.ndc_focal = self.ndc_config.focal
        else:
            opts.ndc_width = -1
        opts.sigma_thresh = 1e-2 if fast else 0.0
        opts.stop_thresh = 1e-2 if fast else 0.0
        if hasattr(self, "sigma_thresh"):
            opts.sigma_thresh = self.sigma_thresh
        if hasattr(self, "stop_thresh"):
            opts.stop_thresh = self.stop_thresh
        return opts
