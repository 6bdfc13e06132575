"""Per-frame rebuild (B200, `pytest -m gpu`): the sort-free bitmap build (svoxb_build_dense.cu) against the sort-based
build (svoxb_build.cu) -- identical tensors -- and the capacity-bounded, synchronisation-free frame path.
Reference semantics: (depth-1) x [query_vertical + N3Tree.refine] + construct_tree (svox_t/svox.py:488-560,
svox_kernel.cu:110-121, 274-324); isomorphism with that loop is covered in test_gpu_parity.py / test_gpu_fullsize.py."""
import numpy as np
import pytest
import torch

import svox_t_b200 as sv
from svox_t_b200 import csrc as C
from svox_t_b200 import synth

pytestmark = pytest.mark.gpu


def _points(L, n, seed, dev, spread=0.45):
    rng = np.random.default_rng(seed)
    p = (0.5 + spread * (rng.random((n, 3)) * 2 - 1)).astype(np.float32)
    p[: n // 8] = p[n // 8: 2 * (n // 8)]                      # exact duplicates: several points per finest cell
    p[-4:] = np.array([[0, 0, 0], [1, 1, 1], [-3, 0.5, 0.5], [0.999999, 0.5, 2.0]], np.float32)   # clamped corners / outside
    return torch.from_numpy(p).to(dev)


@pytest.mark.parametrize("L,n", [(1, 50), (2, 300), (3, 5000), (5, 40000), (8, 300000), (9, 200000), (10, 150000)])
def test_bitmap_build_equals_sort_based_build(dev, L, n):
    pts = _points(L, n, L, dev)
    off, scl = torch.zeros(3, device=dev), torch.ones(3, device=dev)
    a = C.build_octree(pts, L, off, scl)                        # bitmap build, exact size
    b = C.build_octree(pts, L, off, scl, sort_based=True)
    assert a[3] is None and a[0].shape == b[0].shape
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    # non-trivial world -> tree transform
    off2 = torch.tensor([0.1, -0.2, 0.05], device=dev)
    scl2 = torch.tensor([0.8, 1.1, 0.9], device=dev)
    a2, b2 = C.build_octree(pts, L, off2, scl2), C.build_octree(pts, L, off2, scl2, sort_based=True)
    assert all(torch.equal(x, y) for x, y in zip(a2[:3], b2[:3]))


def test_capacity_bounded_build_needs_no_sync(dev):
    L, n = 7, 120000
    pts = _points(L, n, 3, dev, spread=0.3)
    off, scl = torch.zeros(3, device=dev), torch.ones(3, device=dev)
    exact = C.build_octree(pts, L, off, scl)
    need = exact[0].shape[0]
    cap = need + 1000
    child, data, pd, status = C.build_octree(pts, L, off, scl, capacity=cap)
    assert child.shape[0] == cap and status.tolist() == [need, 0]
    assert torch.equal(child[:need], exact[0]) and torch.equal(data[:need], exact[1]) and torch.equal(pd[:need], exact[2])
    assert not bool(child[need:].any()) and bool((data[need:] == 1410065408).all())      # spare rows: unreachable empty nodes
    # too small a capacity: flagged on the device, raised when the node count is asked for
    tree = sv.N3Tree(N=2, data_dim=4, map_location=dev).build_from_points(pts, L, capacity=need - 10)
    with pytest.raises(RuntimeError, match="too small"):
        tree.filled
    # the frame path: rebuild with a capacity, render without ever asking for the node count
    f = torch.from_numpy(synth.synth_features(n, 8)).to(dev)
    o, d = synth.synth_rays(4096)
    rays = sv.Rays(*(torch.from_numpy(x).to(dev) for x in (o, d, d)))
    t_exact = sv.N3Tree(N=2, data_dim=8, map_location=dev).build_from_points(pts, L)
    t_cap = sv.N3Tree(N=2, data_dim=8, map_location=dev).build_from_points(pts, L, capacity=cap)
    out_exact = sv.VolumeRenderer(t_exact)(f, rays)
    out_cap = sv.VolumeRenderer(t_cap)(f, rays)
    assert t_cap.__dict__.get("_filled_pending") is not None       # still unresolved: nothing synchronised
    assert torch.equal(out_exact, out_cap)
    assert t_cap.filled == need == t_exact.filled
    v1 = t_exact(f, pts, want_node_ids=True, want_data_ids=True)
    v2 = t_cap(f, pts, want_node_ids=True, want_data_ids=True)
    assert all(torch.equal(x, y) for x, y in zip(v1, v2))


def test_accelerator_is_refilled_in_place_frame_after_frame(dev):
    """Animated frames: rebuild (capacity-bounded) -> accelerator (refilled in place, svoxb_accel_rebuild) -> render, three
    poses; every frame equals the render of a tree and accelerator built from scratch."""
    L, n = 7, 60000
    f = torch.from_numpy(synth.synth_features(n, 16)).to(dev)
    o, d = synth.synth_rays(8192)
    rays = sv.Rays(*(torch.from_numpy(x).to(dev) for x in (o, d, d)))
    tree = sv.N3Tree(N=2, data_dim=16, map_location=dev)
    r = sv.VolumeRenderer(tree)
    handles = set()
    for seed, spread in ((1, 0.30), (2, 0.22), (3, 0.35)):
        pts = _points(L, n, seed, dev, spread=spread)
        tree.build_from_points(pts, L, capacity=200000)
        out, depth = r.forward_with_depth(f, rays)
        acc = tree.accel(f)
        handles.add(acc.handle.value)
        fresh = sv.N3Tree(N=2, data_dim=16, map_location=dev).build_from_points(pts, L)
        out2, depth2 = sv.VolumeRenderer(fresh).forward_with_depth(f, rays)
        assert torch.equal(out, out2) and torch.equal(depth, depth2)
        assert acc.describe()["bricks"][1] == fresh.accel(f).describe()["bricks"][1]
    assert len(handles) == 1                                   # one accelerator object, refilled


def test_animated_frame_is_cuda_graph_capturable(dev):
    """warp -> splat -> rebuild -> accelerator -> render issue no synchronisation and (after a warm-up frame) allocate
    nothing outside torch's graph pool: the whole frame is captured once and replayed with new joint transforms."""
    L, P, D = 6, 40000, 8
    rng = np.random.default_rng(5)
    pts = (0.5 + 0.25 * (rng.random((P, 3)) * 2 - 1)).astype(np.float32)
    Tm, w, ji = synth.synth_skeleton(P)
    p_t, w_t, j_t = (torch.from_numpy(a).to(dev) for a in (pts, w, ji))
    Tm_t = torch.from_numpy(Tm).to(dev)
    f = torch.from_numpy(synth.synth_features(P, D)).to(dev)
    corner, size = torch.zeros(3, device=dev), torch.ones(3, device=dev)
    cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
    tree = sv.N3Tree(N=2, data_dim=D, map_location=dev)
    r = sv.VolumeRenderer(tree)

    def frame():
        warped, _ = sv.warp_vertices(Tm_t, p_t, w_t, j_t)
        grid = sv.voxelize(warped, f, corner, size, 64, 1.5 / 64, 2.0 / 64)
        tree.build_from_points(warped, L, capacity=80000)
        img, depth = r.render_persp_with_depth(f, cam, width=160, height=120, fx=150.0)
        return grid, img, depth

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                      # warm-up on the capture stream: bitmaps, accelerator, tables
        for _ in range(2):
            frame()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        grid_g, img_g, depth_g = frame()
    # another pose: rotate every joint a little more, replay
    Tm2 = Tm.copy()
    Tm2[:, :3, 3] += 0.01
    Tm_t.copy_(torch.from_numpy(Tm2).to(dev))
    graph.replay()
    torch.cuda.synchronize()
    got = (grid_g.clone(), img_g.clone(), depth_g.clone())
    ref_tree = sv.N3Tree(N=2, data_dim=D, map_location=dev)
    warped, _ = sv.warp_vertices(Tm_t, p_t, w_t, j_t)
    ref_tree.build_from_points(warped, L)
    ref_img, ref_depth = sv.VolumeRenderer(ref_tree).render_persp_with_depth(f, cam, width=160, height=120, fx=150.0)
    assert torch.equal(got[1], ref_img) and torch.equal(got[2], ref_depth)
    assert float((got[0] - sv.voxelize(warped, f, corner, size, 64, 1.5 / 64, 2.0 / 64)).abs().max()) < 1e-4
    assert float((img_g[..., -1] > 0).float().mean()) > 0.05
