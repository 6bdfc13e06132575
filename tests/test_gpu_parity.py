"""GPU parity tests (B200, `pytest -m gpu`): the CUDA path, called through the C-ABI mirror, against
  (1) the reference's own CUDA outputs in tests/golden/*.npz,
  (2) the CPU oracle on the same seeded inputs,
  (3) the compiled reference extension itself when oracle/_ref is present on the box,
and, at BASELINE.json's full sizes, size-independent properties.

Tolerances: leaf indices bit-exact; fwd |a-ref| <= 1e-4 + 1e-3|ref| on >= 99.9 % of
entries and mean abs err <= 1e-5; depth <= 1e-5 on >= 99.9 % of rays; grads relative L2 <= 1e-4.
"""
import os

import numpy as np
import pytest
import torch

import refdrv
import svox_t_b200 as sv
from conftest import GOLDEN_DIR, fmt_case_names, golden_files, golden_fmt_file, golden_variant_files
from oracle import oracle as orc
from svox_t_b200 import csrc as C
from svox_t_b200 import synth

pytestmark = pytest.mark.gpu


def frac_within(a, ref, atol=1e-4, rtol=1e-3):
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    return float((np.abs(a - ref) <= atol + rtol * np.abs(ref)).mean())


def rel_l2(a, ref):
    return float(np.linalg.norm(np.asarray(a, np.float64) - ref) / max(np.linalg.norm(ref), 1e-30))


def make_tree(z_or_tr, D, dev, offset=None, scaling=None):
    t = sv.N3Tree.from_tensors(z_or_tr["child"], z_or_tr["data"], z_or_tr["parent_depth"], data_dim=D,
                               map_location=dev)
    if offset is not None:
        t.offset = torch.from_numpy(np.asarray(offset, np.float32)).to(dev)
        t.invradius = torch.from_numpy(np.asarray(scaling, np.float32)).to(dev)
    return t


def cu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def assert_render_parity(out, depth, grad, ref_out, ref_depth, ref_grad):
    assert frac_within(out, ref_out) >= 0.999
    assert float(np.abs(out - ref_out).mean()) <= 1e-5
    assert float((np.abs(depth - ref_depth) <= 1e-5).mean()) >= 0.999
    assert rel_l2(grad, ref_grad) <= 1e-4


# ---- (1) golden vectors produced by the reference's CUDA kernels ----------------------------------------------------
@pytest.mark.parametrize("accel", [True, False], ids=["accel", "refwalk"])
@pytest.mark.parametrize("name", golden_files())
def test_against_reference_golden(dev, name, accel):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    D = z["features"].shape[1]
    tree = make_tree(z, D, dev, z["offset"], z["scaling"])
    feats = cu(z["features"], dev).requires_grad_(True)
    r = sv.VolumeRenderer(tree)
    r.sigma_thresh, r.stop_thresh = float(z["sigma_thresh"]), float(z["stop_thresh"])
    rays = sv.Rays(cu(z["origins"], dev), cu(z["dirs"], dev), cu(z["dirs"], dev))
    ts = tree._spec(feats, _with_accel=accel)
    assert (ts._accel is not None) == accel
    rs, opt = sv.renderer._rays_spec_from_rays(rays), r._get_options()
    out, depth = C.volume_render_with_depth(ts, rs, opt)
    grad = C.volume_render_backward(ts, rs, opt, cu(z["grad_out"], dev), saved_out=out)
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], grad.cpu().numpy(),
                         z["ref_out"], z["ref_depth"], z["ref_grad"])
    # the pre-activated table (sigmoid once per row) must give the same answers as the in-kernel sigmoid
    ts._act = C.Activated(feats.detach())
    out_a, depth_a = C.volume_render_with_depth(ts, rs, opt)
    grad_a = C.volume_render_backward(ts, rs, opt, cu(z["grad_out"], dev), saved_out=out_a)
    assert_render_parity(out_a.cpu().numpy(), depth_a.cpu().numpy()[:, 0], grad_a.cpu().numpy(),
                         z["ref_out"], z["ref_depth"], z["ref_grad"])
    assert torch.equal(depth_a, depth) and float((out_a - out.detach()).abs().max()) < 1e-6
    ts._act = None
    # standalone depth kernel and the backward without a saved forward agree with the fused path
    assert torch.equal(C.render_depth(ts, rs, opt), depth)
    grad2 = C.volume_render_backward(ts, rs, opt, cu(z["grad_out"], dev))
    assert rel_l2(grad2.cpu().numpy(), z["ref_grad"]) <= 1e-4
    # descent: bit-exact leaf indices, rows and unique-leaf set
    vals, node_ids, data_ids, leaf = C.query_vertical(ts, cu(z["pts"], dev))
    valid = z["ref_valid"]
    assert (node_ids.cpu().numpy() == z["ref_node_ids"]).all()
    assert ((data_ids.cpu().numpy() >= 0) == valid).all()
    assert (data_ids.cpu().numpy()[valid] == z["ref_data_ids"][valid]).all()
    assert (vals.cpu().numpy()[valid] == z["ref_values"][valid]).all()
    assert (leaf.cpu().numpy() == z["ref_leaf_node"]).all()


def test_image_kernel_matches_golden_camera_rays(dev):
    z = np.load(os.path.join(GOLDEN_DIR, "ball_L5_D40_cam.npz"))
    cam = z["camera"]
    c2w, fx, W, H = cam[:16].reshape(4, 4), float(cam[16]), int(cam[18]), int(cam[19])
    D = z["features"].shape[1]
    tree = make_tree(z, D, dev)
    feats = cu(z["features"], dev).requires_grad_(True)
    r = sv.VolumeRenderer(tree)
    img, depth = r.render_persp_with_depth(feats, cu(c2w, dev), width=W, height=H, fx=fx)
    assert img.shape == (H, W, D) and depth.shape == (H, W, 1)
    (img * cu(z["grad_out"], dev).view(H, W, D)).sum().backward()
    assert_render_parity(img.detach().cpu().numpy().reshape(-1, D), depth.cpu().numpy().reshape(-1),
                         feats.grad.cpu().numpy(), z["ref_out"], z["ref_depth"], z["ref_grad"])
    # 3x4 pose and the plain render_persp entry point
    img2 = r.render_persp(feats.detach(), cu(c2w[:3], dev), width=W, height=H, fx=fx)
    assert torch.equal(img2, img.detach())


# ---- (2) CPU oracle on seeded inputs, autograd plumbing, edge cases -------------------------------------------------
@pytest.mark.parametrize("L,shape,D,Q", [(4, "all", 16, 2048), (6, "ball", 33, 2048), (5, "ball", 64, 1024),
                                         (3, "ball", 2, 512), (5, "shell", 100, 512),
                                         (5, "ball", 4, 1024), (5, "ball", 8, 1024), (5, "ball", 12, 1024),
                                         (4, "ball", 128, 512)])
def test_autograd_path_vs_oracle(dev, L, shape, D, Q):
    tr = synth.synth_tree(L, shape, r_out=0.45 if L <= 3 else 0.30, r_in=0.2)
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q)
    g = np.random.default_rng(5).standard_normal((Q, D)).astype(np.float32)
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev).requires_grad_(True)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    out, depth = sv.VolumeRenderer(tree).forward_with_depth(feats, rays)
    (out * cu(g, dev)).sum().backward()
    T = orc.Tree(tr["child"], tr["data"])
    o_ref, d_ref = orc.render_rays(T, f, o, d)
    g_ref = orc.render_rays_backward(T, f, o, d, g)
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                         o_ref, d_ref, g_ref)


def test_non_octree_branching_factor(dev):
    t = sv.N3Tree(N=3, data_dim=5, map_location=dev)
    t.refine(repeats=2)                                   # 27^3 leaves of side 1/27
    n_leaf = t.n_leaves
    leaves = t._all_leaves()
    t.data[(*leaves.T,)] = torch.arange(n_leaf, dtype=torch.int32, device=dev)[:, None]
    t._invalidate()
    f = synth.synth_features(n_leaf, 5)
    o, d = synth.synth_rays(777)
    feats = cu(f, dev).requires_grad_(True)
    assert t.accel(feats) is None                          # packed accelerator is octree-only
    out, depth = sv.VolumeRenderer(t).forward_with_depth(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))
    g = np.random.default_rng(5).standard_normal((777, 5)).astype(np.float32)
    (out * cu(g, dev)).sum().backward()
    T = orc.Tree(t.child.cpu().numpy(), t.data.cpu().numpy())
    o_ref, d_ref = orc.render_rays(T, f, o, d)
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                         o_ref, d_ref, orc.render_rays_backward(T, f, o, d, g))


def test_edge_cases(dev):
    tr = synth.synth_tree(4, "ball")
    D = 8
    tree = make_tree(tr, D, dev)
    f = synth.synth_features(tr["M"], D)
    feats = cu(f, dev)
    r = sv.VolumeRenderer(tree, background_brightness=0.25)
    # empty batch
    e = torch.zeros(0, 3, device=dev)
    assert r(feats, sv.Rays(e, e, e)).shape == (0, D)
    # rays that miss the cube, axis-aligned rays (dir components exactly 0), origin inside the cube, ragged batch size
    o = np.array([[3, 3, 3], [0.5, 0.5, -1], [0.5, 0.5, 0.5], [-1, 0.25, 0.75], [0.5, 2.0, 0.5]], np.float32)
    d = np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0], [1, 0, 0], [0, -1, 0]], np.float32)
    o = np.concatenate([o, synth.synth_rays(32)[0]]); d = np.concatenate([d, synth.synth_rays(32)[1]])
    out, depth = r.forward_with_depth(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))
    T = orc.Tree(tr["child"], tr["data"])
    o_ref, d_ref = orc.render_rays(T, f, o, d, background_brightness=0.25)
    assert np.allclose(out.cpu().numpy(), o_ref, atol=2e-5) and np.allclose(depth.cpu().numpy()[:, 0], d_ref, atol=1e-5)
    assert np.allclose(out[0].cpu().numpy(), [0.25] * (D - 1) + [0.0])       # miss: background, opacity 0
    # all-empty tree (no row anywhere): pure background
    empty = sv.N3Tree(N=2, data_dim=D, init_refine=2, map_location=dev)
    out = sv.VolumeRenderer(empty)(torch.zeros(1, D, device=dev), sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))
    assert torch.allclose(out[:, :-1], torch.ones_like(out[:, :-1])) and not out[:, -1].any()
    # unsupported formats fail loudly instead of rendering something else
    bad = sv.VolumeRenderer(tree)
    bad.data_format = sv.DataFormat("SH9")
    with pytest.raises(RuntimeError):
        bad(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))
    with pytest.raises(RuntimeError):       # float64 features with float32 rays: mixed types are an error, not a cast
        r(feats.double(), sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))


def test_refine_by_points_then_construct_tree(dev):
    """The per-frame rebuild through the reference API: tree[pts].refine() x (L-1), then construct_tree."""
    L = 5
    tr = synth.synth_tree(L, "ball")
    keys = np.nonzero(tr["data"].reshape(-1) != synth.SENTINEL)[0]
    rng = np.random.default_rng(2)
    vox = synth._occupied_keys(L, "ball")
    pts = synth.voxel_centers(vox, L)
    pts = pts[rng.permutation(len(pts))]
    tree = sv.N3Tree(N=2, data_dim=4, init_reserve=64, map_location=dev)
    p = cu(pts, dev)
    for _ in range(L - 1):
        tree[p].refine()
    tree.construct_tree(p)
    assert tree.filled == tr["n_nodes"] and tree.n_leaves == tr["n_leaves"]   # isomorphic to the sorted build
    _, node_ids, data_ids = tree(torch.zeros(len(pts), 4, device=dev), p, want_node_ids=True, want_data_ids=True)
    assert (data_ids.cpu().numpy() == np.arange(len(pts))).all()               # one point per leaf: point i -> row i
    assert len(keys) == len(pts)
    # same geometry => same render as the canonical tree, once rows are permuted accordingly
    f = synth.synth_features(len(pts), 4)
    o, d = synth.synth_rays(512)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    out = sv.VolumeRenderer(tree)(cu(f, dev), rays)
    T = orc.Tree(tree.child.cpu().numpy(), tree.data.cpu().numpy())
    assert frac_within(out.cpu().numpy(), orc.render_rays(T, f, o, d)[0]) >= 0.999


@pytest.mark.parametrize("L", [1, 3, 6])
def test_one_shot_builder_is_isomorphic_to_refine_loop(dev, L):
    rng = np.random.default_rng(2)
    vox = synth._occupied_keys(L, "ball", r_out=0.45 if L < 4 else 0.30)
    centers = synth.voxel_centers(vox, L)
    # several points per voxel, jittered inside it, shuffled
    pts = np.repeat(centers, 3, axis=0) + (rng.random((3 * len(centers), 3)).astype(np.float32) - 0.5) * (0.6 / (1 << L))
    pts = pts[rng.permutation(len(pts))].astype(np.float32)
    p = cu(pts, dev)
    a = sv.N3Tree(N=2, data_dim=4, map_location=dev).build_from_points(p, L)
    b = sv.N3Tree(N=2, data_dim=4, init_reserve=64, map_location=dev)
    for _ in range(L - 1):
        b[p].refine()
    b.construct_tree(p)
    assert a.filled == b.filled and a.n_leaves == b.n_leaves and a.max_depth == b.max_depth == L - 1
    # structural invariants of the emitted tensors
    ch, pd = a.child.reshape(-1, 8).cpu().numpy(), a.parent_depth.cpu().numpy()
    node, slot = np.nonzero(ch)
    kid = node + ch[node, slot]
    assert sorted(kid.tolist()) == list(range(1, a.filled))
    assert (pd[kid, 0] == node * 8 + slot).all() and (pd[kid, 1] == pd[node, 1] + 1).all()
    # same point -> leaf-row map (largest point index of the voxel) and same geometry at arbitrary query points
    f = torch.zeros(len(pts), 4, device=dev)
    _, _, ida = a(f, p, want_node_ids=True, want_data_ids=True)
    _, _, idb = b(f, p, want_node_ids=True, want_data_ids=True)
    assert torch.equal(ida, idb)
    qp = torch.rand(20000, 3, device=dev)
    _, _, qa = a(f, qp, want_node_ids=True, want_data_ids=True)
    _, _, qb = b(f, qp, want_node_ids=True, want_data_ids=True)
    assert torch.equal(qa, qb)
    T = orc.Tree(a.child.cpu().numpy(), a.data.cpu().numpy())
    assert (orc.query(T, np.zeros((len(pts), 4), np.float32), pts)[2] == ida.cpu().numpy()).all()


def test_warp_vertices_and_p2v_vs_oracle(dev):
    P = 5000
    rng = np.random.default_rng(2)
    pts = (0.5 + 0.3 * (rng.random((P, 3)) - 0.5)).astype(np.float32)
    Tm, w, ji = synth.synth_skeleton(P)
    w[:, 3] = 0.0                                            # exercise the w > 0 guard
    co, mats = sv.warp_vertices(cu(Tm, dev), cu(pts, dev), cu(w, dev), cu(ji, dev))
    co_ref, m_ref = orc.warp_vertices(Tm, pts, w, ji)
    assert np.allclose(co.cpu().numpy(), co_ref, atol=1e-6) and np.allclose(mats.cpu().numpy(), m_ref, atol=1e-6)
    assert torch.equal(sv.blend_transformation_matrix(cu(Tm, dev), cu(w, dev), cu(ji, dev)), mats)
    feat = rng.random((P, 3)).astype(np.float32)
    corner, size = np.zeros(3, np.float32), np.ones(3, np.float32)
    vox = sv.voxelize(cu(pts, dev), cu(feat, dev), cu(corner, dev), cu(size, dev), 64, 1.5 / 64, 2.0 / 64)
    v_ref = orc.p2v(pts, feat, corner, size, 64, 1.5 / 64, 2.0 / 64)
    assert vox.shape == (64, 64, 64, 1)
    assert np.allclose(vox.cpu().numpy(), v_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("accel", [True, False], ids=["accel", "refwalk"])
@pytest.mark.parametrize("name", golden_variant_files())
def test_march_variants_against_reference_golden(dev, name, accel):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    D = z["features"].shape[1]
    tree = make_tree(z, D, dev, z["offset"], z["scaling"])
    tree.extra_data = cu(z["extra"], dev)
    feats = cu(z["features"], dev)
    rays = sv.Rays(cu(z["origins"], dev), cu(z["dirs"], dev), cu(z["dirs"], dev))
    rs = sv.renderer._rays_spec_from_rays(rays)
    ts = tree._spec(feats, _with_accel=accel)
    r = sv.VolumeRenderer(tree)
    for tag, fast in (("default", False), ("fast", True)):
        opt = r._get_options(fast)
        op = C.opacity_render(ts, rs, opt).cpu().numpy()[:, 0]
        assert frac_within(op, z["opacity_" + tag]) >= 0.999
        out, dep, hit, idx = (t.cpu().numpy() for t in C.motion_render(ts, rs, opt))
        same = idx[:, 0] == z["motion_idx_" + tag]
        assert same.mean() >= 0.999
        assert np.allclose(dep[same, 0], z["motion_depth_" + tag][same], atol=1e-5)
        assert np.allclose(hit[same], z["motion_hit_" + tag][same], atol=1e-5)
        assert np.allclose(out[same], z["motion_out_" + tag][same], atol=1e-5)
    # opacity backward: the reference never runs its own kernel for it (Appendix B2) -> oracle + autograd plumbing
    fw = feats.clone().requires_grad_(True)
    g = np.random.default_rng(4).standard_normal((len(z["origins"]), 1)).astype(np.float32)
    (r.opacity_render(fw, rays) * cu(g, dev)).sum().backward()
    T = orc.Tree(z["child"], z["data"], z["offset"], z["scaling"])
    g_ref = orc.opacity_render_backward(T, z["features"], z["origins"], z["dirs"], g)
    assert rel_l2(fw.grad.cpu().numpy(), g_ref) <= 1e-4
    # the opacity channel of the full render and the opacity-only render agree
    full = r(feats, rays)
    assert torch.allclose(full[:, -1:], C.opacity_render(ts, rs, r._get_options()), atol=1e-6)


def test_point_kernels_forward_backward_against_reference_golden(dev):
    z = np.load(os.path.join(GOLDEN_DIR, "x_points_lbs_p2v.npz"))
    T = cu(z["T"], dev).requires_grad_(True)
    x = cu(z["pts"], dev).requires_grad_(True)
    w = cu(z["w"], dev).requires_grad_(True)
    co, mats = sv.warp_vertices(T, x, w, cu(z["ji"], dev))
    assert np.allclose(co.detach().cpu().numpy(), z["ref_coords"], atol=1e-6)
    assert np.allclose(mats.detach().cpu().numpy(), z["ref_mats"], atol=1e-6)
    ((co * cu(z["g_coords"], dev)).sum() + (mats * cu(z["g_mats"], dev)).sum()).backward()
    assert np.allclose(x.grad.cpu().numpy(), z["ref_gx"], atol=1e-5)
    assert np.allclose(w.grad.cpu().numpy(), z["ref_gw"], atol=1e-5)
    assert rel_l2(T.grad.cpu().numpy(), z["ref_gT"]) <= 1e-4
    n, kr, cr = int(z["n_voxels"]), float(z["kernel_radius"]), float(z["conv_radius"])
    p = cu(z["ref_coords"], dev).requires_grad_(True)
    f = cu(z["feat"], dev).requires_grad_(True)
    vox = sv.voxelize(p, f, cu(z["corner"], dev), cu(z["size"], dev), n, kr, cr)
    assert np.allclose(vox.detach().cpu().numpy(), z["ref_vox"], rtol=1e-4, atol=1e-4)
    (vox * cu(z["g_vox"], dev)).sum().backward()
    assert np.allclose(f.grad.cpu().numpy(), z["ref_gf"], rtol=1e-4, atol=1e-4)
    assert np.allclose(p.grad.cpu().numpy(), z["ref_gp"], rtol=1e-3, atol=1e-2)
    # many joints: the bone-gradient table no longer fits shared memory -> global-atomic path
    P, J = 4000, 5000
    rng = np.random.default_rng(1)
    Tb = rng.standard_normal((J, 4, 4)).astype(np.float32)
    wb = rng.dirichlet(np.ones(3), P).astype(np.float32)
    jb = rng.integers(0, J, (P, 3)).astype(np.int32)
    xb = rng.random((P, 3)).astype(np.float32)
    gc, gm = rng.standard_normal((P, 3)).astype(np.float32), rng.standard_normal((P, 4, 4)).astype(np.float32)
    gx, gT, gw = C.warp_vertices_backward(cu(Tb, dev), cu(xb, dev), cu(wb, dev), cu(jb, dev), cu(gc, dev), cu(gm, dev))
    ox, oT, ow = orc.warp_vertices_backward(Tb, xb, wb, jb, gc, gm)
    assert np.allclose(gx.cpu().numpy(), ox, atol=1e-4) and np.allclose(gw.cpu().numpy(), ow, atol=1e-4)
    assert rel_l2(gT.cpu().numpy(), oT) <= 1e-5


# ---- (3) the compiled reference itself, when it travelled to this box -----------------------------------------------
@pytest.mark.skipif(not os.path.exists(refdrv.REF_SO), reason="oracle/_ref not built")
def test_against_live_reference_extension(dev):
    m = refdrv.module()
    tr = synth.synth_tree(7, "ball")
    D, Q = 32, 20000
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q, seed=11)
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev).requires_grad_(True)
    o_t, d_t = cu(o, dev), cu(d, dev)
    g_t = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    out, depth = sv.VolumeRenderer(tree).forward_with_depth(feats, sv.Rays(o_t, d_t, d_t))
    (out * g_t).sum().backward()
    rts = refdrv.tree_spec(feats.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius,
                           tree.filled)
    rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
    assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                         m.volume_render(rts, rrs, ro).cpu().numpy(), m.render_depth(rts, rrs, ro).cpu().numpy()[:, 0],
                         m.volume_render_backward(rts, rrs, ro, g_t).cpu().numpy())
    pts = torch.rand(100000, 3, device=dev) * 1.2 - 0.1
    v, nid, did, leaf = tree(feats.detach(), pts, want_node_ids=True, want_data_ids=True, want_leaf_node=True)
    rv, rnid, rdid, rleaf = m.query_vertical(rts, pts)
    ok = did >= 0
    assert torch.equal(nid, rnid) and torch.equal(did[ok], rdid[ok]) and torch.equal(v[ok], rv[ok])
    assert torch.equal(leaf, rleaf[torch.argsort(tree._pack_index(rleaf))])


# ---- full-size properties (BASELINE configs C2 / C3) -----------------------------------------------------------------
@pytest.fixture(scope="module")
def c3_scene(dev):
    tr = synth.synth_tree(8, "ball")
    D = 32
    tree = make_tree(tr, D, dev)
    feats = cu(synth.synth_features(tr["M"], D), dev)
    return tr, tree, feats


def test_full_size_c3_properties(dev, c3_scene):
    tr, tree, feats = c3_scene
    assert tr["M"] == 1897408 and tr["n_nodes"] == 281697
    Q = 1 << 20
    o, d = synth.synth_rays(Q)
    o_t, d_t = cu(o, dev), cu(d, dev)
    r = sv.VolumeRenderer(tree)
    fw = feats.clone().requires_grad_(True)
    out, depth = r.forward_with_depth(fw, sv.Rays(o_t, d_t, d_t))
    assert torch.isfinite(out).all()
    op = out[:, -1]
    assert float(op.min()) >= 0 and float(op.max()) <= 1 and float((op > 0).float().mean()) > 0.95
    assert float(out[:, :-1].min()) >= 0 and float(out[:, :-1].max()) <= 1 + 1e-5   # convex mix of sigmoids and bg
    # invariance to ray order (ray compaction must not mix rays up)
    perm = torch.randperm(Q, device=dev)
    out_p = r(feats, sv.Rays(o_t[perm].contiguous(), d_t[perm].contiguous(), d_t[perm].contiguous()))
    assert torch.equal(out_p, out.detach()[perm])
    # linearity of the backward in grad_output and additivity over ray subsets
    g1, g2 = torch.randn(Q, 32, device=dev), torch.randn(Q, 32, device=dev)
    ts = tree._spec(feats)
    rs, opt = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t)), r._get_options()
    G1 = C.volume_render_backward(ts, rs, opt, g1, saved_out=out.detach())
    G2 = C.volume_render_backward(ts, rs, opt, g2, saved_out=out.detach())
    G12 = C.volume_render_backward(ts, rs, opt, (g1 + 2 * g2).contiguous(), saved_out=out.detach())
    assert float((G12 - (G1 + 2 * G2)).norm() / G12.norm()) < 1e-4
    h = Q // 2
    rsa = sv.renderer._rays_spec_from_rays(sv.Rays(o_t[:h], d_t[:h], d_t[:h]))
    rsb = sv.renderer._rays_spec_from_rays(sv.Rays(o_t[h:], d_t[h:], d_t[h:]))
    Ga = C.volume_render_backward(ts, rsa, opt, g1[:h].contiguous(), saved_out=out.detach()[:h].contiguous())
    Gb = C.volume_render_backward(ts, rsb, opt, g1[h:].contiguous(), saved_out=out.detach()[h:].contiguous())
    assert float((Ga + Gb - G1).norm() / G1.norm()) < 1e-5
    # a sample of the full batch against the oracle
    T = orc.Tree(tr["child"], tr["data"])
    sel = np.arange(0, Q, Q // 2048)[:2048]
    o_ref, d_ref = orc.render_rays(T, feats.cpu().numpy(), o[sel], d[sel])
    assert frac_within(out.detach().cpu().numpy()[sel], o_ref) >= 0.999
    assert float((np.abs(depth.cpu().numpy()[sel, 0] - d_ref) <= 1e-5).mean()) >= 0.999


def test_full_size_c2_image(dev, c3_scene):
    tr, tree, feats = c3_scene
    r = sv.VolumeRenderer(tree)
    cam = cu(synth.synth_cameras(1)[0], dev)
    img, depth = r.render_persp_with_depth(feats, cam)                      # 800 x 800, fx = 1111.111 (API defaults)
    assert img.shape == (800, 800, 32)
    hit = img[..., -1] > 0
    assert 0.55 < float(hit.float().mean()) < 0.66                          # ~61 % of pixels hit
    assert bool(((depth[..., 0] > 0) == (img[..., -1] > 0)).float().mean() > 0.99)
    # the same pixels rendered as explicit rays (restated cam2world_ray) agree statistically
    oc, dc = orc.camera_rays(cam.cpu().numpy(), 1111.111, 1111.111, 800, 800)
    out = r(feats, sv.Rays(cu(oc, dev), cu(dc, dev), cu(dc, dev)))
    assert frac_within(out.cpu().numpy(), img.cpu().numpy().reshape(-1, 32)) >= 0.999


# ---- view-dependent formats, NDC cameras, motion-feature render -------------------------------------
_FMT_NAMES = {1: "SH", 2: "SG", 3: "ASG"}


def _fmt_tree(z, D, fmt, B, extra, dev, accel=True):
    tree = sv.N3Tree.from_tensors(z["child"], z["data"], z["parent_depth"], data_dim=D,
                                  data_format=f"{_FMT_NAMES[fmt]}{B}", map_location=dev)
    if extra is not None:
        tree.extra_data = cu(extra, dev)
    if not accel:
        tree.accel = lambda *a, **k: None          # walk the reference tensors instead of the packed accelerator
    return tree


@pytest.mark.parametrize("accel", [True, False], ids=["accel", "refwalk"])
def test_view_dependent_formats_against_reference_golden(dev, accel):
    z = np.load(golden_fmt_file())
    T = orc.Tree(z["child"], z["data"])
    rays = sv.Rays(cu(z["origins"], dev), cu(z["dirs"], dev), cu(z["vdirs"], dev))
    for name in fmt_case_names(z):
        fmt, B, Cc, cmin, cmax, with_tm = (int(v) for v in z[name + "_meta"])
        f, g, thr = z[name + "_features"], z[name + "_grad_out"], float(z[name + "_thresh"])
        extra = z[name + "_extra"] if name + "_extra" in z.files else None
        tree = _fmt_tree(z, f.shape[1], fmt, B, extra, dev, accel)
        r = sv.VolumeRenderer(tree, min_comp=cmin, max_comp=cmax)
        r.sigma_thresh, r.stop_thresh = thr, thr
        feats = cu(f, dev).requires_grad_(True)
        out = r(feats, rays, transformation_matrices=cu(z["tm"], dev) if with_tm else None)
        assert tuple(out.shape) == (len(z["origins"]), Cc + 1)
        (out * cu(g, dev)).sum().backward()
        out_n = out.detach().cpu().numpy()
        assert frac_within(out_n, z[name + "_ref_out"]) >= 0.999, name
        assert float(np.abs(out_n - z[name + "_ref_out"]).mean()) <= 1e-5, name
        if with_tm:     # the reference's gradient uses a stale basis here: compare with the oracle's correct one
            ref_grad = orc.render_rays_fmt_backward(T, f, z["origins"], z["dirs"], z["vdirs"], g, fmt, B, extra=extra,
                                                    tm=z["tm"], min_comp=cmin, max_comp=cmax)
        else:
            ref_grad = z[name + "_ref_grad"]
        assert rel_l2(feats.grad.cpu().numpy(), ref_grad) <= 1e-4, name


def test_motion_feature_render_forward_reference_backward_oracle(dev):
    """Both forms of the motion-feature render: the table form (large batches: Q * 32 >= M, blend + sigmoid once per row,
    the feature render's quad kernels on that table) and the staged kernels (small batches)."""
    z = np.load(golden_fmt_file())
    f4, jf, sw, ji = z["mf_features"], z["mf_jf"], z["mf_sw"], z["mf_ji"]
    T = orc.Tree(z["child"], z["data"])
    tree = make_tree(z, 4, dev)
    M = f4.shape[0]
    for nq in (M // 32 - 1, len(z["origins"])):                    # staged kernels / table form
        o_np, d_np = z["origins"][:nq], z["dirs"][:nq]
        rays = sv.Rays(cu(o_np, dev), cu(d_np, dev), cu(d_np, dev))
        for tag, thr in (("default", 0.0), ("fast", 1e-2)):
            r = sv.VolumeRenderer(tree, background_brightness=0.5)
            r.sigma_thresh, r.stop_thresh = thr, thr
            jft = cu(jf, dev).requires_grad_(True)
            out = r.motion_feature_render(cu(f4, dev), jft, cu(sw, dev), cu(ji, dev), rays)
            assert frac_within(out.detach().cpu().numpy(), z["mf_ref_out_" + tag][:nq]) >= 0.999
            g = np.random.default_rng(1).standard_normal(tuple(out.shape)).astype(np.float32)
            (out * cu(g, dev)).sum().backward()
            g_ref = orc.motion_feature_render_backward(T, f4, o_np, d_np, jf, sw, ji, g)
            assert rel_l2(jft.grad.cpu().numpy(), g_ref) <= 1e-4
        # B != 4 (scalar staging), with a negative weight (skipped, rt_kernel.cu:955) and a repeated joint per row
        sw3, ji3 = np.ascontiguousarray(sw[:, :3]).copy(), np.ascontiguousarray(ji[:, :3]).copy()
        sw3[::7, 1] = -0.25
        ji3[::5, 2] = ji3[::5, 0]
        for accel in (True, False):
            ts = tree._spec(cu(f4, dev), cu(jf, dev), cu(sw3, dev), cu(ji3, dev), _with_accel=accel)
            rs, opt = sv.renderer._rays_spec_from_rays(rays), sv.VolumeRenderer(tree)._get_options()
            out = C.motion_feature_render(ts, rs, opt)
            g = np.random.default_rng(2).standard_normal(tuple(out.shape)).astype(np.float32)
            gj = C.motion_feature_render_backward(ts, rs, opt, cu(g, dev))
            assert frac_within(out.cpu().numpy(), orc.motion_feature_render(T, f4, o_np, d_np, jf, sw3, ji3)) >= 0.999
            assert rel_l2(gj.cpu().numpy(), orc.motion_feature_render_backward(T, f4, o_np, d_np, jf, sw3, ji3, g)) <= 1e-4
        # feature widths on either side of the float4 layouts: F + 1 a multiple of 4 (full rows) and not
        for F2 in (3, 7, 9):
            jf2 = np.random.default_rng(F2).standard_normal((jf.shape[0], F2)).astype(np.float32)
            ts = tree._spec(cu(f4, dev), cu(jf2, dev), cu(sw, dev), cu(ji, dev))
            out = C.motion_feature_render(ts, rs, opt)
            g = np.random.default_rng(4).standard_normal(tuple(out.shape)).astype(np.float32)
            gj = C.motion_feature_render_backward(ts, rs, opt, cu(g, dev))
            assert frac_within(out.cpu().numpy(), orc.motion_feature_render(T, f4, o_np, d_np, jf2, sw, ji)) >= 0.999
            assert rel_l2(gj.cpu().numpy(), orc.motion_feature_render_backward(T, f4, o_np, d_np, jf2, sw, ji, g)) <= 1e-4
    rays = sv.Rays(cu(z["origins"], dev), cu(z["dirs"], dev), cu(z["dirs"], dev))
    # rays that miss the cube return zeros, not the background (rt_kernel.cu:911-916)
    o = np.array([[3, 3, 3], [0.5, 0.5, -1.0]], np.float32)
    d = np.array([[0, 1, 0], [0, 0, 1]], np.float32)
    out = sv.VolumeRenderer(tree, background_brightness=0.5).motion_feature_render(
        cu(f4, dev), cu(jf, dev), cu(sw, dev), cu(ji, dev), sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)))
    assert not out[0].any() and out[1].any()
    o2, d2 = np.concatenate([z["origins"], o]), np.concatenate([z["dirs"], d])          # the same two in a table-form batch
    out = sv.VolumeRenderer(tree, background_brightness=0.5).motion_feature_render(
        cu(f4, dev), cu(jf, dev), cu(sw, dev), cu(ji, dev), sv.Rays(cu(o2, dev), cu(d2, dev), cu(d2, dev)))
    assert not out[-2].any() and out[-1].any()
    # many joints: the per-CTA gradient table no longer fits shared memory -> global atomics
    J2 = 2000
    rng = np.random.default_rng(3)
    jf2 = rng.standard_normal((J2, 16)).astype(np.float32)
    ji2 = rng.integers(0, J2, ji.shape).astype(np.int32)
    jft = cu(jf2, dev).requires_grad_(True)
    out = sv.VolumeRenderer(tree).motion_feature_render(cu(f4, dev), jft, cu(sw, dev), cu(ji2, dev), rays)
    g = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    (out * cu(g, dev)).sum().backward()
    assert frac_within(out.detach().cpu().numpy(), orc.motion_feature_render(T, f4, z["origins"], z["dirs"], jf2, sw, ji2)) >= 0.999
    assert rel_l2(jft.grad.cpu().numpy(),
                  orc.motion_feature_render_backward(T, f4, z["origins"], z["dirs"], jf2, sw, ji2, g)) <= 1e-4


@pytest.mark.parametrize("ndc", [False, True], ids=["world", "ndc"])
def test_image_render_sh_and_ndc_vs_oracle(dev, ndc):
    """Camera rays generated in-kernel (+ NDC conversion, rt_kernel.cu:1168-1206) for an SH tree and for the RGBA
    format, against the oracle marching the oracle's own camera rays. The reference cannot run this path (fact #4)."""
    tr = synth.synth_tree(5, "ball")
    T = orc.Tree(tr["child"], tr["data"])
    W, H, fx = 52, 37, 60.0
    if ndc:    # forward-facing set-up: camera in front of the NDC cube looking down -z
        c2w = synth.look_at((0.1, -0.05, 3.0), target=(0.0, 0.0, 0.0))
        nd = sv.NDCConfig(W, H, fx)
        kw = dict(ndc_width=W, ndc_height=H, ndc_focal=fx)
        radius, center = 1.0, 0.0
    else:
        c2w = synth.synth_cameras(1)[0]
        nd, kw, radius, center = None, {}, 0.5, 0.5
    inv, off = np.full(3, 0.5 / radius), np.full(3, 0.5 * (1.0 - center / radius))
    T = orc.Tree(tr["child"], tr["data"], off, inv)
    o, d, vd = orc.camera_rays_ndc(c2w, fx, fx, W, H, **kw)
    g_rng = np.random.default_rng(8)
    for fmtname, D in (("SH4", 13), ("RGBA", 8)):
        f = synth.synth_features(tr["M"], D)
        tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, radius=radius,
                                      center=[center] * 3, data_format=fmtname, map_location=dev)
        r = sv.VolumeRenderer(tree, ndc=nd)
        feats = cu(f, dev).requires_grad_(True)
        img = r.render_persp(feats, cu(c2w, dev), width=W, height=H, fx=fx)
        Do = img.shape[-1]
        g = g_rng.standard_normal((H * W, Do)).astype(np.float32)
        (img * cu(g, dev).view(H, W, Do)).sum().backward()
        if fmtname == "RGBA":
            o_ref = orc.render_rays(T, f, o, d)[0]
            g_ref = orc.render_rays_backward(T, f, o, d, g)
        else:
            o_ref = orc.render_rays_fmt(T, f, o, d, vd, orc.FORMAT_SH, 4)
            g_ref = orc.render_rays_fmt_backward(T, f, o, d, vd, g, orc.FORMAT_SH, 4)
        assert o_ref[:, -1].max() > 0.5                                   # the object is in view
        assert frac_within(img.detach().cpu().numpy().reshape(-1, Do), o_ref) >= 0.999
        assert rel_l2(feats.grad.cpu().numpy(), g_ref) <= 1e-4


def test_accumulate_weights_context(dev):
    """`with tree.accumulate_weights() as accum:` around ray and image renders (svox.py:664-677, rt_kernel.cu:308-310)."""
    tr = synth.synth_tree(5, "ball")
    D = 8
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(3000)
    tree = make_tree(tr, D, dev)
    r = sv.VolumeRenderer(tree)
    feats = cu(f, dev)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    with tree.accumulate_weights() as accum:
        with pytest.raises(RuntimeError):
            tree.refine()                                        # structure is locked while weights accumulate
        out = r(feats, rays)
        per_leaf = accum()
        raw = accum.value.clone()
    assert tree._weight_accum is None
    T = orc.Tree(tr["child"], tr["data"])
    ref = orc.accumulate_weights(T, f, o, d)
    assert raw.shape == tree.child.shape and per_leaf.shape[0] == tree.n_leaves
    assert np.allclose(raw.cpu().numpy(), ref, rtol=1e-4, atol=1e-5)
    assert abs(float(raw.sum()) - float(out[:, -1].sum())) <= 1e-3 * float(out[:, -1].sum())
    # image render accumulates too; a second render inside the same block adds on top
    cam = cu(synth.synth_cameras(1)[0], dev)
    with tree.accumulate_weights() as accum:
        img = r.render_persp(feats, cam, width=64, height=48, fx=70.0)
        once = float(accum.value.sum())
        r.render_persp(feats, cam, width=64, height=48, fx=70.0)
        twice = float(accum.value.sum())
    assert abs(once - float(img[..., -1].sum())) <= 1e-3 * once and abs(twice - 2 * once) <= 1e-3 * once


@pytest.mark.parametrize("D", [16, 13], ids=["row_kernels", "scalar_lane_kernels"])
def test_hit_marks_are_an_exact_acceleration(dev, D):
    """Rows marked sigma <= 0 are skipped without being fetched: outputs and gradients do not change, stale marks are
    not trusted after an in-place update of the features, and a negative sigma_thresh ignores them. (D = 13 without an
    activated table takes the scalar-lane kernels, whose samples honour the marks too.)"""
    tr = synth.synth_tree(6, "ball")
    Q = 4096
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q)
    g = cu(np.random.default_rng(2).standard_normal((Q, D)).astype(np.float32), dev)
    tree = make_tree(tr, D, dev)
    r = sv.VolumeRenderer(tree)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    rs, opt = sv.renderer._rays_spec_from_rays(rays), r._get_options()
    feats = cu(f, dev)
    ts_plain = tree._spec(feats)                                  # accelerator, no marks
    ref_out = C.volume_render(ts_plain, rs, opt)
    ref_grad = C.volume_render_backward(ts_plain, rs, opt, g, saved_out=ref_out)
    ts = tree._spec(feats)
    ts._accel.mark_hits(feats)
    assert ts._c().accel_marks_current == 1
    out = C.volume_render(ts, rs, opt)
    assert torch.equal(out, ref_out)
    grad = C.volume_render_backward(ts, rs, opt, g, saved_out=out)
    assert float((grad - ref_grad).norm() / ref_grad.norm()) < 1e-6
    # flip the sign of every sigma in place: the marks are now wrong, and must no longer be used
    feats[:, -1].neg_()
    assert ts._c().accel_marks_current == 0
    flipped = C.volume_render(ts, rs, opt)
    T = orc.Tree(tr["child"], tr["data"])
    f2 = f.copy()
    f2[:, -1] *= -1
    assert frac_within(flipped.cpu().numpy()[:512], orc.render_rays(T, f2, o[:512], d[:512])[0]) >= 0.999
    # a threshold below zero admits rows with thresh < sigma <= 0, which the marks (sigma > 0) would drop: the forward
    # must ignore them then. (Empty leaves count as sigma = 0 in the reference and would dereference a null row for
    # such a threshold, rt_kernel.cu:278-304, so only rows that exist are compared: here against the unmarked run.)
    ts._accel.mark_hits(feats)
    r.sigma_thresh = -1.0
    lo = C.volume_render(ts, rs, r._get_options())
    assert torch.equal(lo, C.volume_render(ts_plain, rs, r._get_options()))
    assert not torch.equal(lo, flipped)


def test_deep_tree_three_stage_accelerator(dev):
    """Depth-9 shell: the accelerator needs three stages ([4,3,2]; the third lookup happens after the compositing of the
    previous sample). Packed walk == reference walk bit for bit; a ray sample against the oracle; gradients too."""
    tr = synth.synth_tree(9, "shell", r_out=0.30, r_in=0.2985)
    D, Q = 16, 20000
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q)
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev)
    acc = tree.accel(feats)
    assert acc.describe()["stages"] == 3 and tree.max_depth == 8
    r = sv.VolumeRenderer(tree)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    rs, opt = sv.renderer._rays_spec_from_rays(rays), r._get_options()
    out_a, dep_a = C.volume_render_with_depth(tree._spec(feats), rs, opt)
    out_r, dep_r = C.volume_render_with_depth(tree._spec(feats, _with_accel=False), rs, opt)
    assert torch.equal(dep_a, dep_r) and float((out_a - out_r).abs().max()) <= 1e-6
    g = cu(np.random.default_rng(3).standard_normal((Q, D)).astype(np.float32), dev)
    ga = C.volume_render_backward(tree._spec(feats), rs, opt, g, saved_out=out_a)
    gr = C.volume_render_backward(tree._spec(feats, _with_accel=False), rs, opt, g, saved_out=out_r)
    assert float((ga - gr).norm() / gr.norm()) <= 1e-5
    T = orc.Tree(tr["child"], tr["data"])
    n = 1024
    o_ref, d_ref = orc.render_rays(T, f, o[:n], d[:n])
    assert frac_within(out_a.cpu().numpy()[:n], o_ref) >= 0.999
    assert float((np.abs(dep_a.cpu().numpy()[:n, 0] - d_ref) <= 1e-5).mean()) >= 0.999
    g_ref = orc.render_rays_backward(T, f, o[:n], d[:n], g.cpu().numpy()[:n])
    rs_n = sv.renderer._rays_spec_from_rays(sv.Rays(cu(o[:n], dev), cu(d[:n], dev), cu(d[:n], dev)))
    g_n = C.volume_render_backward(tree._spec(feats), rs_n, opt, g[:n].contiguous(), saved_out=out_a[:n].contiguous())
    assert rel_l2(g_n.cpu().numpy(), g_ref) <= 1e-4
    # points query on the deep tree: leaf ids bit-exact with the oracle
    pts = np.random.default_rng(4).random((5000, 3)).astype(np.float32)
    _, nid, did = tree(feats, cu(pts, dev), want_node_ids=True, want_data_ids=True)
    _, on, od, ov = orc.query(T, f, pts)
    assert (nid.cpu().numpy() == on).all() and (did.cpu().numpy()[ov] == od[ov]).all()


@pytest.mark.parametrize("B", [1, 4, 9, 16, 25])
def test_sh_rgb_fast_path_vs_oracle(dev, B):
    """SH rows with three output channels take the lane-private kernels (svoxb_render_shrgb.cu): every degree, with
    and without the packed accelerator, per-row rotations, a component window, thresholds, rays and camera images."""
    tr = synth.synth_tree(5, "ball")
    T = orc.Tree(tr["child"], tr["data"])
    D, Q = 3 * B + 1, 1500
    f = synth.synth_features(tr["M"], D, seed=B)
    f[:, :-1] *= 0.6
    rng = np.random.default_rng(B)
    o, d = synth.synth_rays(Q, seed=2)
    vd = (synth._unit(rng, Q) * rng.uniform(0.5, 1.5, (Q, 1))).astype(np.float32)
    tm = np.zeros((tr["M"], 4, 4), np.float32)
    for i in range(tr["M"]):
        tm[i, :3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    g = rng.standard_normal((Q, 4)).astype(np.float32)
    lo, hi = (1, B - 2) if B >= 4 else (0, B - 1)
    for accel in (True, False):
        for with_tm, (cmin, cmax), thr in ((False, (0, B - 1), 0.0), (True, (lo, hi), 0.0), (False, (0, B - 1), 1e-2)):
            tree = _fmt_tree(tr, D, 1, B, None, dev, accel)
            r = sv.VolumeRenderer(tree, min_comp=cmin, max_comp=cmax)
            r.sigma_thresh, r.stop_thresh = thr, thr
            feats = cu(f, dev).requires_grad_(True)
            out = r(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(vd, dev)),
                    transformation_matrices=cu(tm, dev) if with_tm else None)
            (out * cu(g, dev)).sum().backward()
            kw = dict(tm=tm if with_tm else None, min_comp=cmin, max_comp=cmax)
            o_ref = orc.render_rays_fmt(T, f, o, d, vd, orc.FORMAT_SH, B, sigma_thresh=thr, stop_thresh=thr, **kw)
            g_ref = orc.render_rays_fmt_backward(T, f, o, d, vd, g, orc.FORMAT_SH, B, **kw)
            assert frac_within(out.detach().cpu().numpy(), o_ref) >= 0.999, (B, accel, with_tm)
            assert rel_l2(feats.grad.cpu().numpy(), g_ref) <= 1e-4, (B, accel, with_tm)
    # camera image through the same kernels
    W, H, fx = 40, 30, 45.0
    c2w = synth.synth_cameras(1)[0]
    tree = _fmt_tree(tr, D, 1, B, None, dev)
    feats = cu(f, dev).requires_grad_(True)
    img = sv.VolumeRenderer(tree).render_persp(feats, cu(c2w, dev), width=W, height=H, fx=fx)
    gi = rng.standard_normal((H * W, 4)).astype(np.float32)
    (img * cu(gi, dev).view(H, W, 4)).sum().backward()
    oc, dc, vc = orc.camera_rays_ndc(c2w, fx, fx, W, H)
    assert frac_within(img.detach().cpu().numpy().reshape(-1, 4), orc.render_rays_fmt(T, f, oc, dc, vc, orc.FORMAT_SH, B)) >= 0.999
    assert rel_l2(feats.grad.cpu().numpy(), orc.render_rays_fmt_backward(T, f, oc, dc, vc, gi, orc.FORMAT_SH, B)) <= 1e-4


def test_image_bands_tile_the_full_frame(dev):
    """rows=(y0, y1) renders a band of the frame: bands tile the full image bit for bit (outputs and depth), and the
    gradient of a band loss is the gradient of the same loss on those rows of the full frame."""
    from svox_t_b200 import dist as svd
    tr = synth.synth_tree(5, "ball")
    D, W, H, fx = 8, 70, 53, 80.0
    tree = make_tree(tr, D, dev)
    feats = cu(synth.synth_features(tr["M"], D), dev)
    cam = cu(synth.synth_cameras(1)[0], dev)
    r = sv.VolumeRenderer(tree)
    full = r.render_persp(feats, cam, width=W, height=H, fx=fx)
    bounds = [svd.shard_image_rows(H, k, 3) for k in range(3)]
    assert bounds[0][0] == 0 and bounds[-1][1] == H and all(b[0] % 8 == 0 for b in bounds)
    bands = [r.render_persp(feats, cam, width=W, height=H, fx=fx, rows=b) for b in bounds]
    assert torch.equal(torch.cat(bands, 0), full)
    assert torch.equal(svd.render_image_bands(r, feats, cam, W, H, fx), full)           # one rank: the whole frame
    y0, y1 = bounds[1]
    g = torch.randn(y1 - y0, W, D, device=dev)
    fa = feats.clone().requires_grad_(True)
    (r.render_persp(fa, cam, width=W, height=H, fx=fx, rows=(y0, y1)) * g).sum().backward()
    fb = feats.clone().requires_grad_(True)
    gfull = torch.zeros(H, W, D, device=dev)
    gfull[y0:y1] = g
    (r.render_persp(fb, cam, width=W, height=H, fx=fx) * gfull).sum().backward()
    assert float((fa.grad - fb.grad).norm() / fb.grad.norm()) < 1e-5
    with pytest.raises(RuntimeError):
        r.render_persp(feats, cam, width=W, height=H, fx=fx, rows=(8, H + 1))


@pytest.mark.parametrize("D", [2, 3, 5, 6, 7, 9, 13, 17, 31, 33, 63, 65, 127, 128])
def test_every_width_rays_and_images_vs_oracle(dev, D):
    """Sweep of feature widths through the large-batch path (activated table, hit marks; D % 4 != 0: payload-only table,
    compact sigma, scratch gradients): explicit rays and camera images, default and `fast` thresholds, depth."""
    tr = synth.synth_tree(5, "ball")
    T = orc.Tree(tr["child"], tr["data"])
    f = synth.synth_features(tr["M"], D, seed=D)
    tree = make_tree(tr, D, dev)
    r = sv.VolumeRenderer(tree, background_brightness=0.4)
    rng = np.random.default_rng(D)
    Q = 2048                                                   # Q * 32 >= M: the renderer attaches the derived tables
    o, d = synth.synth_rays(Q, seed=D)
    g = rng.standard_normal((Q, D)).astype(np.float32)
    for fast in (False, True):
        thr = 1e-2 if fast else 0.0
        feats = cu(f, dev).requires_grad_(True)
        out, depth = r.forward_with_depth(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev)), fast=fast)
        (out * cu(g, dev)).sum().backward()
        o_ref, d_ref = orc.render_rays(T, f, o, d, background_brightness=0.4, sigma_thresh=thr, stop_thresh=thr)
        g_ref = orc.render_rays_backward(T, f, o, d, g, background_brightness=0.4)
        assert_render_parity(out.detach().cpu().numpy(), depth.cpu().numpy()[:, 0], feats.grad.cpu().numpy(),
                             o_ref, d_ref, g_ref)
    W, H, fx = 61, 43, 70.0
    c2w = synth.synth_cameras(1)[0]
    oc, dc = orc.camera_rays(c2w, fx, fx, W, H)
    gi = rng.standard_normal((H * W, D)).astype(np.float32)
    feats = cu(f, dev).requires_grad_(True)
    img, dep = r.render_persp_with_depth(feats, cu(c2w, dev), width=W, height=H, fx=fx)
    (img * cu(gi, dev).view(H, W, D)).sum().backward()
    o_ref, d_ref = orc.render_rays(T, f, oc, dc, background_brightness=0.4)
    assert_render_parity(img.detach().cpu().numpy().reshape(-1, D), dep.cpu().numpy().reshape(-1),
                         feats.grad.cpu().numpy(), o_ref, d_ref,
                         orc.render_rays_backward(T, f, oc, dc, gi, background_brightness=0.4))


def test_snap_and_clone_on_device(dev):
    tr = synth.synth_tree(4, "ball")
    tree = make_tree(tr, 4, dev)
    pts = torch.rand(500, 3, device=dev)
    corners = tree.snap(pts)
    view = tree[pts]
    lengths = view.lengths[torch.searchsorted(tree._pack_index(view.unique_leaf_node), view.leaf_node_id)]
    assert corners.shape == (500, 3)
    assert bool(((pts >= corners - 1e-6) & (pts < corners + lengths + 1e-6)).all())
    c2 = tree.clone()
    assert c2.child.is_cuda and torch.equal(c2.child, tree.child) and c2.child.data_ptr() != tree.child.data_ptr()
    cpu = tree.clone(device="cpu")
    assert not cpu.child.is_cuda and torch.equal(cpu.data, tree.data.cpu())


# ---- the remaining point-wise operators of svox_t.csrc (svoxb_vertical.cu) -----------------------------------------
def test_query_backward_and_assign_vs_oracle(dev):
    """query_vertical_backward / assign_vertical (svox_kernel.cu:83-108): the reference's kernels fault (Appendix B1),
    so parity is against the oracle's restatement of their source and against torch index arithmetic."""
    tr = synth.synth_tree(5, "ball")
    D, Q = 19, 5000
    f = synth.synth_features(tr["M"], D)
    T = orc.Tree(tr["child"], tr["data"])
    rng = np.random.default_rng(21)
    pts = (rng.random((Q, 3)) * 1.1 - 0.05).astype(np.float32)          # some outside the cube, many in empty leaves
    tree = make_tree(tr, D, dev)
    feats = cu(f, dev).requires_grad_(True)
    vals = tree(feats, cu(pts, dev))
    g = rng.standard_normal((Q, D)).astype(np.float32)
    n0 = C.launch_count()
    (vals * cu(g, dev)).sum().backward()
    assert C.launch_count() > n0                                          # the backward is one of our kernels
    ref = orc.query_backward(T, tr["M"], pts, g, dtype=np.float64)
    assert ref.any() and rel_l2(feats.grad.cpu().numpy(), ref) <= 1e-6
    # K != D through the operator module directly
    g5 = np.ascontiguousarray(g[:, :5])
    got = C.query_vertical_backward(tree._spec(feats.detach()), cu(pts, dev), cu(g5, dev)).cpu().numpy()
    assert got.shape == (tr["M"], 5) and rel_l2(got, orc.query_backward(T, tr["M"], pts, g5, dtype=np.float64)) <= 1e-6

    # assignment: whole rows, and the leading K channels only; duplicates resolved towards the largest point index
    pts2 = np.concatenate([pts, pts[:700]])                                # 700 leaves are hit (at least) twice
    for K in (D, 4):
        v = rng.standard_normal((len(pts2), K)).astype(np.float32)
        tree.features = torch.nn.Parameter(cu(f, dev))
        ver = tree.features._version
        tree.set(cu(pts2, dev), cu(v, dev))
        assert tree.features._version > ver
        assert np.array_equal(tree.features.detach().cpu().numpy(), orc.assign(T, f, pts2, v))
    tree.features = torch.nn.Parameter(cu(f, dev))
    tree[cu(pts2, dev)] = cu(v, dev)
    assert np.array_equal(tree.features.detach().cpu().numpy(), orc.assign(T, f, pts2, v))
    # empty batches are no-ops
    tree.set(torch.zeros((0, 3), device=dev), torch.zeros((0, D), device=dev))
    assert C.query_vertical_backward(tree._spec(feats.detach()), torch.zeros((0, 3), device=dev),
                                     torch.zeros((0, D), device=dev)).abs().sum().item() == 0.0


@pytest.mark.parametrize("N", [2, 3])
def test_calc_corners_vs_oracle_and_query(dev, N):
    """calc_corners (svox_kernel.cu:213-237; raises in the reference, Appendix B4): bit-exact against the oracle, and
    every corner + half a cell must query back into the same leaf."""
    tree = sv.N3Tree(N=N, data_dim=4, init_reserve=64, map_location=dev)
    gen = torch.Generator(device="cpu").manual_seed(3)
    for _ in range(3):
        pts = torch.rand(40, 3, generator=gen).to(dev)
        tree[pts].refine()
    view = tree[:]
    leaves = view.unique_leaf_node.contiguous()
    got = C.calc_corners(tree._spec(tree.features), leaves)
    ref = orc.calc_corners(tree.parent_depth.cpu().numpy(), N, leaves.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), ref)
    assert torch.equal(view.corners_local, got)
    centre = got + 0.5 * view.lengths_local[:, None]
    _, nid = tree(tree.features, tree.tree2world(centre).contiguous(), want_node_ids=True)
    assert torch.equal(nid, tree._pack_index(leaves))
    assert C.calc_corners(tree._spec(tree.features), leaves[:0]).shape == (0, 3)


def _grid_case(ndc):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_grid", os.path.join(GOLDEN_DIR, "make_golden_grid.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.grid_case(ndc)


@pytest.mark.parametrize("ndc", [False, True], ids=["world", "ndc"])
def test_grid_weight_render_vs_oracle_and_reference(dev, ndc):
    """grid_weight_render (rt_kernel.cu:1240-1344): hit counts bit-exact but for threshold ties, maximum weights within
    the float tolerance; against the oracle and, when oracle/_ref is on the box, the reference's own kernel."""
    grid, c2w, W, H, fx, off, inv, kw = _grid_case(ndc)
    cam = C.CameraSpec()
    cam.c2w, cam.fx, cam.fy, cam.width, cam.height = cu(c2w, dev), fx, fx, W, H
    opt = C.RenderOptions()
    opt.step_size, opt.sigma_thresh = 1e-3, 0.5
    for k, v in kw.items():
        setattr(opt, k, v)
    gw, gh = C.grid_weight_render(cu(grid, dev), cam, opt, cu(off, dev), cu(inv, dev))
    gw, gh = gw.cpu().numpy(), gh.cpu().numpy()
    rw, rh = orc.grid_weight_render(grid, c2w, fx, fx, W, H, off, inv, step_size=1e-3, sigma_thresh=0.5, **kw)
    assert rh.sum() > 1000 and (gh[grid <= 0.5] == 0).all()
    assert float((gh == rh).mean()) >= 0.999 and abs(gh.sum() - rh.sum()) <= 1e-3 * rh.sum()
    assert frac_within(gw, rw, atol=1e-5) >= 0.999
    if refdrv.available():
        m = refdrv.module()
        rc, ro = m.CameraSpec(), refdrv.options(sigma_thresh=0.5)
        rc.c2w, rc.fx, rc.fy, rc.width, rc.height = cu(c2w, dev), fx, fx, W, H
        for k, v in kw.items():
            setattr(ro, k, v)
        xw, xh = m.grid_weight_render(cu(grid, dev), rc, ro, cu(off, dev), cu(inv, dev))
        assert float((gh == xh.cpu().numpy()).mean()) >= 0.999
        assert frac_within(gw, xw.cpu().numpy(), atol=1e-5) >= 0.999


def test_view_values_and_set(dev):
    """N3TreeView.values / .set on feature rows (the reference's accessors still index `data` as floats)."""
    tr = synth.synth_tree(4, "ball")
    D = 6
    f = synth.synth_features(tr["M"], D)
    tree = make_tree(tr, D, dev)
    tree.features = torch.nn.Parameter(cu(f, dev))
    pts = torch.rand(300, 3, device=dev)
    view = tree[pts]
    rows = tree.data[view.key][..., 0].long()
    ok = rows < tr["M"]
    assert 0 < int(ok.sum()) < len(view)
    vals = view.values
    assert torch.equal(vals[ok], tree.features[rows[ok]]) and float(vals[~ok].abs().sum()) == 0.0
    vals.sum().backward()
    want = torch.zeros(tr["M"], device=dev)
    want[rows[ok]] = 1.0
    assert torch.equal(tree.features.grad[:, 0], want)
    view.set(torch.full((len(view), D), 7.0, device=dev))
    assert float(tree.features.detach()[rows[ok]].min()) == 7.0
    assert int((tree.features.detach() == 7.0).all(dim=1).sum()) == int(ok.sum())
    view.clamp_(min=8.0)
    assert float(tree.features.detach()[rows[ok]].min()) == 8.0 and view.shape == (len(view), D) and view.ndim == 2
    view.sigmoid_(); view.relu_(); view.nan_to_num_(); view.uniform_(2.0, 3.0)
    got = tree.features.detach()[rows[ok]]
    assert float(got.min()) >= 2.0 and float(got.max()) <= 3.0 and "leaves" in repr(view)
    untouched = torch.ones(tr["M"], dtype=torch.bool, device=dev)
    untouched[rows[ok]] = False
    assert torch.equal(tree.features.detach()[untouched], cu(f, dev)[untouched])
    tree[pts[:5]] = 3.0                                           # scalar broadcast through assign_vertical
    got = tree(tree.features.detach(), pts[:5], want_data_ids=True)
    assert bool(((got[0] == 3.0).all(dim=1) | (got[1] < 0)).all())


def test_training_step_is_cuda_graph_capturable(dev):
    """The forward + autograd backward issue no synchronisation and allocate only stream-ordered memory, so a step can be
    captured once and replayed (new rays are copied into the captured buffers)."""
    tr = synth.synth_tree(4, "ball")
    D, Q = 16, 2048
    tree = make_tree(tr, D, dev)
    feats = cu(synth.synth_features(tr["M"], D), dev).requires_grad_(True)
    o, d = synth.synth_rays(Q)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    r = sv.VolumeRenderer(tree)
    g = torch.randn(Q, D, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # warm-up off the default stream: accelerator, activated table
        for _ in range(2):
            feats.grad = None
            r(feats, rays).backward(g)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    feats.grad = None
    with torch.cuda.graph(graph):
        out_g = r(feats, rays)
        out_g.backward(g)
    o2, d2 = synth.synth_rays(Q, seed=9)
    rays.origins.copy_(cu(o2, dev)); rays.dirs.copy_(cu(d2, dev))
    graph.replay()
    torch.cuda.synchronize()
    got_out, got_grad = out_g.detach().clone(), feats.grad.clone()
    feats.grad = None
    ref = r(feats, rays)
    ref.backward(g)
    assert torch.equal(got_out, ref.detach())
    assert float((got_grad - feats.grad).norm() / feats.grad.norm()) < 1e-5


@pytest.mark.parametrize("fmt,B", [(2, 4), (2, 9), (3, 4), (3, 16)], ids=["SG4", "SG9", "ASG4", "ASG16"])
def test_sg_asg_rgb_fast_path_vs_oracle(dev, fmt, B):
    """SG / ASG rows with three output channels and B in {1, 4, 9, 16, 25} share the lane-private kernels of the SH-RGB
    layout (only the per-ray basis differs, rt_kernel.cu:116-140): with and without per-row rotations, both tree walks."""
    tr = synth.synth_tree(5, "ball")
    T = orc.Tree(tr["child"], tr["data"])
    D, Q = 3 * B + 1, 1200
    rng = np.random.default_rng(10 * fmt + B)
    f = synth.synth_features(tr["M"], D, seed=B)
    f[:, :-1] *= 0.6
    o, d = synth.synth_rays(Q, seed=3)
    vd = synth._unit(rng, Q).astype(np.float32)
    if fmt == 2:
        extra = np.concatenate([rng.uniform(1, 5, (B, 1)), synth._unit(rng, B)], 1).astype(np.float32)
    else:
        frames = np.stack([np.linalg.qr(rng.standard_normal((3, 3)))[0] for _ in range(B)])       # rows x, y, z
        extra = np.concatenate([rng.uniform(0.5, 3, (B, 2)), frames.reshape(B, 9)], 1).astype(np.float32)
    tm = np.zeros((tr["M"], 4, 4), np.float32)
    for i in range(tr["M"]):
        tm[i, :3, :3] = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    g = rng.standard_normal((Q, 4)).astype(np.float32)
    for accel in (True, False):
        for with_tm in (False, True):
            tree = _fmt_tree(tr, D, fmt, B, extra, dev, accel)
            feats = cu(f, dev).requires_grad_(True)
            out = sv.VolumeRenderer(tree)(feats, sv.Rays(cu(o, dev), cu(d, dev), cu(vd, dev)),
                                          transformation_matrices=cu(tm, dev) if with_tm else None)
            (out * cu(g, dev)).sum().backward()
            kw = dict(extra=extra, tm=tm if with_tm else None)
            o_ref = orc.render_rays_fmt(T, f, o, d, vd, fmt, B, **kw)
            g_ref = orc.render_rays_fmt_backward(T, f, o, d, vd, g, fmt, B, **kw)
            assert np.abs(o_ref[:, :3]).max() > 0.05
            assert frac_within(out.detach().cpu().numpy(), o_ref) >= 0.999, (accel, with_tm)
            assert rel_l2(feats.grad.cpu().numpy(), g_ref) <= 1e-4, (accel, with_tm)


def test_forward_step_counts_order_the_backward(dev):
    """Batches of at least svoxb_ray_order_min_rays() rays: the forward also writes every ray's march-iteration count, the
    backward marches the rays longest first by it. The counts are exact (they equal the oracle's sample counters);
    outputs are those of the unordered calls, gradients too up to the order of the floating-point reductions."""
    lib = C.load_library()
    tr = synth.synth_tree(6, "ball")
    D, Q = 32, 131072
    assert lib.svoxb_ray_order_min_rays() <= Q <= lib.svoxb_ray_order_max_rays()
    tree = make_tree(tr, D, dev)
    f = synth.synth_features(tr["M"], D)
    feats = cu(f, dev)
    o, d = synth.synth_rays(Q, seed=21)
    o_t, d_t = cu(o, dev), cu(d, dev)
    g = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
    r = sv.VolumeRenderer(tree)
    ts = r._render_spec(feats, Q)
    rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
    opt = r._get_options()
    out = C.volume_render(ts, rs, opt)                             # leaves rs._cost = march iterations per ray
    assert rs._cost is not None and int(rs._cost.min()) >= 0
    T = orc.Tree(tr["child"], tr["data"])
    sel = np.arange(0, Q, 64)
    cnt = orc.render_rays(T, f, o[sel], d[sel], want_counters=True)[2]
    assert int(rs._cost.cpu().numpy()[sel].sum()) == int(cnt["S"])             # exact: the same sample sequence
    grad = C.volume_render_backward(ts, rs, opt, g, saved_out=out)
    # the plain C-ABI calls (no counts, no order)
    out2 = torch.empty_like(out)
    C._check(lib.svoxb_render_rays_fwd(C.ctypes.byref(ts._c()), C._ptr(o_t), C._ptr(d_t), C._ptr(d_t), Q,
                                       C.ctypes.byref(opt._c()), C._ptr(out2), None, C._stream()))
    grad2 = torch.zeros_like(feats)
    C._check(lib.svoxb_render_rays_bwd(C.ctypes.byref(ts._c()), C._ptr(o_t), C._ptr(d_t), C._ptr(d_t), Q,
                                       C.ctypes.byref(opt._c(sigma_thresh=0.0, stop_thresh=-1.0)), C._ptr(g), C._ptr(out2),
                                       C._ptr(grad2), C._stream()))
    assert torch.equal(out, out2)
    assert float((grad - grad2).norm() / grad2.norm()) < 1e-6
    # a forward that cannot count (fused depth) says so; the backward then keeps the caller's order
    rs2 = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
    out3, _ = C.volume_render_with_depth(ts, rs2, opt)
    assert int(rs2._cost[0]) == -1
    grad3 = C.volume_render_backward(ts, rs2, opt, g, saved_out=out3)
    assert float((grad3 - grad2).norm() / grad2.norm()) < 1e-6
    # a large batch (2^19 rays) is ordered too
    Q2 = 1 << 19
    o2, d2 = synth.synth_rays(Q2, seed=22)
    o2_t, d2_t = cu(o2, dev), cu(d2, dev)
    g2 = torch.randn(Q2, D, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
    rs3 = sv.renderer._rays_spec_from_rays(sv.Rays(o2_t, d2_t, d2_t))
    out4 = C.volume_render(ts, rs3, opt)
    assert rs3._cost is not None
    grad4 = C.volume_render_backward(ts, rs3, opt, g2, saved_out=out4)
    grad5 = torch.zeros_like(feats)
    C._check(lib.svoxb_render_rays_bwd(C.ctypes.byref(ts._c()), C._ptr(o2_t), C._ptr(d2_t), C._ptr(d2_t), Q2,
                                       C.ctypes.byref(opt._c(sigma_thresh=0.0, stop_thresh=-1.0)), C._ptr(g2), C._ptr(out4),
                                       C._ptr(grad5), C._stream()))
    assert float((grad4 - grad5).norm() / grad5.norm()) < 1e-6


def test_fresh_feature_tensors_never_meet_a_stale_table(dev):
    """Features are typically fresh network outputs every step: version 0, and the caching allocator hands the freed
    block back at the same address. The activated table and the hit marks are keyed on the storage OBJECT, so the
    second tensor must not be rendered (or back-propagated) with the first one's tables."""
    tr = synth.synth_tree(5, "ball")
    D, Q = 16, 20000
    M = tr["M"]
    tree = make_tree(tr, D, dev)
    o, d = synth.synth_rays(Q, seed=4)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    rs, r = sv.renderer._rays_spec_from_rays(rays), sv.VolumeRenderer(tree)
    opt = r._get_options()
    g = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    ptrs = set()
    for seed in range(5):
        gen = torch.Generator(device=dev).manual_seed(100 + seed)
        f = torch.randn(M, D, device=dev, generator=gen)
        f[:, -1] = torch.rand(M, device=dev, generator=gen) * 10 - (2 + seed)      # another set of dead rows every time
        ptrs.add(f.data_ptr())
        f.requires_grad_(True)
        out = r(f, rays)                                       # cached tables: activated + hit marks
        (out * g).sum().backward()
        plain = tree._spec(f.detach(), _with_accel=False)      # no accelerator, no activated table, no marks
        ref_out = C.volume_render(plain, rs, opt)
        ref_grad = C.volume_render_backward(plain, rs, opt, g, saved_out=ref_out)
        assert float((out.detach() - ref_out).abs().max()) <= 2e-6
        assert float((f.grad - ref_grad).norm() / ref_grad.norm()) <= 1e-5
        del f, out, plain, ref_out, ref_grad
    assert len(ptrs) < 5                                       # the allocator did hand an address back at least once


@pytest.mark.parametrize("known_depth", [False, True], ids=["accel_sync", "accel_no_readback"])
def test_rows_shared_by_several_leaves_keep_exact_hit_marks(dev, known_depth):
    """After refine() the children of a split leaf inherit its row (svox.py:539-540): one row, eight leaf cells. The
    table pass cannot mark through its row -> cell map then and must fall back to the pass over the cells -- decided on
    the host when the accelerator was built with read-backs, on the device when it was built without."""
    tr = synth.synth_tree(4, "ball")
    D, Q = 16, 30000
    tree = make_tree(tr, D, dev)
    tree.refine()                                            # every leaf splits once; rows are now shared 8-fold
    f = synth.synth_features(tr["M"], D)
    f[:, -1] = np.random.default_rng(3).uniform(-4, 6, tr["M"]).astype(np.float32)     # ~40 % dead rows
    feats = cu(f, dev)
    o, d = synth.synth_rays(Q, seed=8)
    rays = sv.Rays(cu(o, dev), cu(d, dev), cu(d, dev))
    rs = sv.renderer._rays_spec_from_rays(rays)
    r = sv.VolumeRenderer(tree)
    opt = r._get_options()
    if known_depth:
        acc = tree.accel(feats, max_depth=tree.max_depth + 1)
        assert acc is not None
    g = torch.randn(Q, D, device=dev, generator=torch.Generator(device=dev).manual_seed(2))
    ts = r._render_spec(feats, Q)                            # activated table + hit marks
    assert ts._accel is not None and ts._act is not None
    out = C.volume_render(ts, rs, opt)
    grad = C.volume_render_backward(ts, rs, opt, g, saved_out=out)
    plain = tree._spec(feats, _with_accel=False)
    out_ref = C.volume_render(plain, rs, opt)
    grad_ref = C.volume_render_backward(plain, rs, opt, g, saved_out=out_ref)
    assert float((out - out_ref).abs().max()) <= 2e-6
    assert float((grad - grad_ref).norm() / grad_ref.norm()) <= 1e-5
    T = orc.Tree(tree.child[:tree.filled].cpu().numpy(), tree.data[:tree.filled].cpu().numpy())
    o_ref = orc.render_rays(T, f, o[:2000], d[:2000])[0]
    assert frac_within(out.cpu().numpy()[:2000], o_ref) >= 0.999
