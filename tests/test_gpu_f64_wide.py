"""GPU parity tests of the GENERAL march kernels (svoxb_render_wide.cu), called through the C-ABI mirror:

* the float64 instantiation of the hot path (the reference dispatches AT_DISPATCH_FLOATING_TYPES on every entry point,
  rt_kernel.cu:1373, svox_kernel.cu:290) against the fp64 oracle and, when oracle/_ref travelled to the box, against
  the compiled reference run in double;
* feature widths above 128 in float32 (the reference loops over any out_data_dim, rt_kernel.cu:302-306) against the
  oracle, the compiled reference, and -- channel slice by channel slice -- the tuned D <= 128 kernels.

Tolerances: fp64 vs the fp64 oracle 1e-10 (same arithmetic, only FMA contraction differs); fp64 vs the reference's double
build 1e-5 (that build evaluates exp() in float, rt_kernel.cu:280,304); float32 as in test_gpu_parity.py.
"""
import os

import numpy as np
import pytest
import torch

import refdrv
import svox_t_b200 as sv
from oracle import oracle as orc
from svox_t_b200 import csrc as C
from svox_t_b200 import synth
from test_gpu_parity import assert_render_parity, cu, make_tree, rel_l2

pytestmark = pytest.mark.gpu

STEP = float(np.float32(1e-3))      # the option fields are float in the reference too: the kernels see this value
OFFSET, SCALING = (0.5, 0.4375, 0.5625), (0.375, 0.5, 0.4375)      # exact in float32 (N3Tree keeps them in float32)


def _scene(L, D, Q, seed=3):
    tr = synth.synth_tree(L, "ball")
    f = synth.synth_features(tr["M"], D, seed=seed).astype(np.float64)
    o, d = synth.synth_rays(Q, seed=seed + 1)
    # world space: undo the tree transform so the rays still cross the ball
    o = ((o.astype(np.float64) - np.asarray(OFFSET)) / np.asarray(SCALING))
    d = d.astype(np.float64) / np.asarray(SCALING)
    g = np.random.default_rng(seed + 2).standard_normal((Q, D))
    return tr, f, o, d, g


def _spec64(tr, f, dev):
    tree = make_tree(tr, f.shape[1], dev, OFFSET, SCALING)
    feats = cu(f, dev)
    assert feats.dtype == torch.float64
    ts = tree._spec(feats)
    assert ts.is_f64 and ts._accel is None and ts.offset.dtype == torch.float64
    return tree, feats, ts


@pytest.mark.parametrize("thresholds", [(0.0, 0.0), (0.5, 0.02)], ids=["default", "fast"])
def test_float64_render_matches_fp64_oracle(dev, thresholds):
    sigma_thresh, stop_thresh = thresholds
    tr, f, o, d, g = _scene(5, 9, 6000)
    tree, feats, ts = _spec64(tr, f, dev)
    rs = C.RaysSpec()
    rs.origins, rs.dirs, rs.vdirs = cu(o, dev), cu(d, dev), cu(d, dev)
    opt = C.RenderOptions()
    opt.background_brightness, opt.sigma_thresh, opt.stop_thresh = 0.25, sigma_thresh, stop_thresh
    out, depth = C.volume_render_with_depth(ts, rs, opt)
    assert out.dtype == torch.float64 and depth.dtype == torch.float64
    T = orc.Tree(tr["child"], tr["data"], OFFSET, SCALING)
    o_ref, d_ref = orc.render_rays(T, f, o, d, step_size=STEP, background_brightness=0.25, sigma_thresh=sigma_thresh,
                                   stop_thresh=stop_thresh, dtype=np.float64)[:2]
    assert np.abs(out.cpu().numpy() - o_ref).max() <= 1e-10
    assert np.abs(depth.cpu().numpy()[:, 0] - d_ref).max() <= 1e-10
    assert torch.equal(C.render_depth(ts, rs, opt), depth)
    # backward: with and without the saved forward (non-default thresholds force the re-render)
    g_ref = orc.render_rays_backward(T, f, o, d, g, step_size=STEP, background_brightness=0.25, dtype=np.float64)
    for saved in (out, None):
        grad = C.volume_render_backward(ts, rs, opt, cu(g, dev), saved_out=saved)
        assert grad.dtype == torch.float64 and rel_l2(grad.cpu().numpy(), g_ref) <= 1e-10
    # point query in double
    pts = np.random.default_rng(9).uniform(-1.4, 1.4, (20000, 3))
    v, nid, did, leaf = C.query_vertical(ts, cu(pts, dev))
    rv, rnid, rdid, rvalid = orc.query(T, f, pts, dtype=np.float64)
    rvalid = rvalid.astype(bool)
    assert (nid.cpu().numpy() == rnid).all() and ((did.cpu().numpy() >= 0) == rvalid).all()
    assert (did.cpu().numpy()[rvalid] == rdid[rvalid]).all() and (v.cpu().numpy()[rvalid] == rv[rvalid]).all()
    assert (leaf.cpu().numpy() == orc.leafset(rnid, 2)).all()


def test_float64_through_the_renderer_and_autograd(dev):
    """VolumeRenderer / N3Tree with a double feature table: forward, depth, autograd backward, camera render."""
    tr, f, o, d, g = _scene(4, 6, 3000)
    tree = make_tree(tr, 6, dev, OFFSET, SCALING)
    feats = cu(f, dev).requires_grad_(True)
    r = sv.VolumeRenderer(tree)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    out, depth = r.forward_with_depth(feats, rays)
    (out * cu(g, dev)).sum().backward()
    T = orc.Tree(tr["child"], tr["data"], OFFSET, SCALING)
    o_ref, d_ref = orc.render_rays(T, f, o, d, step_size=STEP, dtype=np.float64)[:2]
    g_ref = orc.render_rays_backward(T, f, o, d, g, step_size=STEP, dtype=np.float64)
    assert np.abs(out.detach().cpu().numpy() - o_ref).max() <= 1e-10
    assert np.abs(depth.cpu().numpy()[:, 0] - d_ref).max() <= 1e-10
    assert rel_l2(feats.grad.cpu().numpy(), g_ref) <= 1e-10
    # camera rays generated in the kernel == the oracle's camera rays rendered as a batch; band == slice of the image
    W, H, fx = 96, 64, 80.0
    eye = (np.asarray([0.5, 0.5, 0.5]) + 1.1 * np.asarray([0.3, 0.5, 0.81]) - np.asarray(OFFSET)) / np.asarray(SCALING)
    c2w = synth.look_at(eye, target=(np.asarray([0.5, 0.5, 0.5]) - np.asarray(OFFSET)) / np.asarray(SCALING)).astype(np.float64)
    feats2 = cu(f, dev).requires_grad_(True)
    img, dimg = r.render_persp_with_depth(feats2, cu(c2w, dev), width=W, height=H, fx=fx)
    assert img.dtype == torch.float64 and img.shape == (H, W, 6)
    gi = np.random.default_rng(4).standard_normal((H * W, 6))
    (img * cu(gi, dev).view(H, W, 6)).sum().backward()
    co, cd = orc.camera_rays(c2w, fx, fx, W, H, dtype=np.float64)
    i_ref, di_ref = orc.render_rays(T, f, co, cd, step_size=STEP, dtype=np.float64)[:2]
    assert np.abs(img.detach().cpu().numpy().reshape(-1, 6) - i_ref).max() <= 1e-10
    assert np.abs(dimg.cpu().numpy().reshape(-1) - di_ref).max() <= 1e-10
    assert rel_l2(feats2.grad.cpu().numpy(), orc.render_rays_backward(T, f, co, cd, gi, step_size=STEP,
                                                                      dtype=np.float64)) <= 1e-10
    cam = sv.renderer._make_camera_spec(cu(c2w, dev), W, H, fx, fx, rows=(16, 40))
    band = C.volume_render_image(tree._spec(feats2.detach()), cam, r._get_options())
    assert torch.equal(band, img.detach()[16:40])
    # mixed types are an error, not a silent cast
    with pytest.raises(RuntimeError):
        r(feats, sv.Rays(cu(o.astype(np.float32), dev), cu(d.astype(np.float32), dev), cu(d.astype(np.float32), dev)))
    # the float32 operators that have no double instantiation say so
    with pytest.raises(RuntimeError):
        r.opacity_render(feats.detach(), rays)


@pytest.mark.skipif(not os.path.exists(refdrv.REF_SO), reason="oracle/_ref not built")
def test_float64_against_the_reference_run_in_double(dev):
    m = refdrv.module()
    tr, f, o, d, g = _scene(6, 16, 20000)
    tree, feats, ts = _spec64(tr, f, dev)
    rs = C.RaysSpec()
    rs.origins, rs.dirs, rs.vdirs = cu(o, dev), cu(d, dev), cu(d, dev)
    opt = C.RenderOptions()
    out, depth = C.volume_render_with_depth(ts, rs, opt)
    grad = C.volume_render_backward(ts, rs, opt, cu(g, dev), saved_out=out)
    rts = refdrv.tree_spec(feats, tree.child, tree.data, tree.parent_depth, ts.offset, ts.scaling, tree.filled,
                           dtype=torch.float64)
    rrs, ro = refdrv.rays_spec(rs.origins, rs.dirs, dtype=torch.float64), refdrv.options()
    r_out, r_depth = m.volume_render(rts, rrs, ro), m.render_depth(rts, rrs, ro)
    r_grad = m.volume_render_backward(rts, rrs, ro, cu(g, dev))
    assert r_out.dtype == torch.float64
    # the reference's double build calls expf(): float-accurate exponentials, everything else in double
    assert float((out - r_out).abs().max()) <= 1e-5
    assert float((depth - r_depth).abs().max()) <= 1e-9
    assert rel_l2(grad.cpu().numpy(), r_grad.cpu().numpy()) <= 1e-5
    pts = torch.rand(50000, 3, device=dev, dtype=torch.float64) * 3.0 - 1.5
    v, nid, did, leaf = C.query_vertical(ts, pts)
    rv, rnid, rdid, rleaf = m.query_vertical(rts, pts)
    ok = did >= 0
    assert torch.equal(nid, rnid) and torch.equal(did[ok], rdid[ok]) and torch.equal(v[ok], rv[ok])
    assert torch.equal(leaf, rleaf[torch.argsort(tree._pack_index(rleaf))])


@pytest.mark.parametrize("D", [129, 200, 260])
def test_wide_feature_rows_float32(dev, D):
    """D > 128: the general kernels, against the oracle and against the tuned kernels run on channel slices."""
    L, Q = 5, 5000
    tr = synth.synth_tree(L, "ball")
    f = synth.synth_features(tr["M"], D, seed=7)
    o, d = synth.synth_rays(Q, seed=8)
    g = np.random.default_rng(10).standard_normal((Q, D)).astype(np.float32)
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev).requires_grad_(True)
    r = sv.VolumeRenderer(tree)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    out, depth = r.forward_with_depth(feats, rays)
    (out * cu(g, dev)).sum().backward()
    T = orc.Tree(tr["child"], tr["data"])
    o_ref, d_ref = orc.render_rays(T, f, o, d)[:2]
    g_ref = orc.render_rays_backward(T, f, o, d, g)
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                         o_ref, d_ref, g_ref)
    # channel slices [c0, c1) + sigma through the tuned kernels: the payload channels composite independently
    for c0, c1 in ((0, 31), (100, D - 1)):
        cols = list(range(c0, c1)) + [D - 1]
        fs = cu(np.ascontiguousarray(f[:, cols]), dev)
        ts = make_tree(tr, len(cols), dev)
        part = sv.VolumeRenderer(ts)(fs, rays)
        assert float((part[:, :-1] - out.detach()[:, c0:c1]).abs().max()) <= 2e-6
        assert float((part[:, -1] - out.detach()[:, -1]).abs().max()) <= 2e-6
    # fast thresholds (early stop + rescale) and the image entry point
    r.sigma_thresh, r.stop_thresh = 0.5, 0.02
    fast = r(feats.detach(), rays)
    f_ref = orc.render_rays(T, f, o, d, sigma_thresh=0.5, stop_thresh=0.02)[0]
    assert np.abs(fast.cpu().numpy() - f_ref).max() <= 2e-5
    r.sigma_thresh, r.stop_thresh = 0.0, 0.0
    W, H, fx = 64, 48, 60.0
    c2w = synth.synth_cameras(1, dist=1.1)[0]
    img = r.render_persp(feats.detach(), cu(c2w, dev), width=W, height=H, fx=fx)
    co, cd = orc.camera_rays(c2w, fx, fx, W, H)
    assert np.abs(img.cpu().numpy().reshape(-1, D) - orc.render_rays(T, f, co, cd)[0]).max() <= 2e-5


@pytest.mark.skipif(not os.path.exists(refdrv.REF_SO), reason="oracle/_ref not built")
def test_wide_feature_rows_against_live_reference(dev):
    m = refdrv.module()
    tr = synth.synth_tree(6, "ball")
    D, Q = 160, 8000
    f = synth.synth_features(tr["M"], D, seed=12)
    o, d = synth.synth_rays(Q, seed=13)
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev).requires_grad_(True)
    o_t, d_t = cu(o, dev), cu(d, dev)
    g_t = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
    out, depth = sv.VolumeRenderer(tree).forward_with_depth(feats, sv.Rays(o_t, d_t, d_t))
    (out * g_t).sum().backward()
    rts = refdrv.tree_spec(feats.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius,
                           tree.filled)
    rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                         m.volume_render(rts, rrs, ro).cpu().numpy(), m.render_depth(rts, rrs, ro).cpu().numpy()[:, 0],
                         m.volume_render_backward(rts, rrs, ro, g_t).cpu().numpy())
