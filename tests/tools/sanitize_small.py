"""Small end-to-end run for compute-sanitizer: every kernel family once on tiny inputs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
for (L, shape, D, accel) in [(4, "ball", 32, True), (3, "ball", 16, False), (4, "ball", 64, True), (3, "ball", 9, True),
                             (5, "shell", 100, True), (4, "ball", 33, True), (4, "ball", 2, True), (4, "ball", 127, False)]:
    tr = synth.synth_tree(L, shape, r_out=0.45, r_in=0.2)
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(2100)          # > M / 32: the renderer attaches the activated table and the hit marks
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    tree.extra_data = torch.rand(4, 3, device=dev)
    feats = torch.from_numpy(f).to(dev).requires_grad_(True)
    rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
    r = sv.VolumeRenderer(tree)
    if not accel:
        tree.accel = lambda *a, **k: None
    out, depth = r.forward_with_depth(feats, rays)
    out.sum().backward()
    img, dep = r.render_persp_with_depth(feats.detach(), torch.from_numpy(synth.synth_cameras(1)[0]).to(dev), width=37, height=21, fx=30.0)
    img2 = r.render_persp(feats, torch.from_numpy(synth.synth_cameras(1)[0]).to(dev), width=37, height=21, fx=30.0)
    img2.sum().backward()
    r.render_depth(feats.detach(), rays); r.opacity_render(feats, rays).sum().backward(); r.motion_render(feats.detach(), rays)
    tree(feats.detach(), torch.rand(500, 3, device=dev), want_node_ids=True, want_leaf_node=True)
pts = torch.rand(3000, 3, device=dev) * 0.5 + 0.25
t2 = sv.N3Tree(N=2, data_dim=8, map_location=dev).build_from_points(pts, 5)
t3 = sv.N3Tree(N=2, data_dim=8, init_reserve=8, map_location=dev)
for _ in range(3): t3[pts].refine()
t3.construct_tree(pts)
Tm, w, ji = synth.synth_skeleton(3000)
wv, mats = sv.warp_vertices(torch.from_numpy(Tm).to(dev), pts, torch.from_numpy(w).to(dev), torch.from_numpy(ji).to(dev))
sv.voxelize(wv, torch.rand(3000, 4, device=dev), torch.zeros(3, device=dev), torch.ones(3, device=dev), 32, 0.05, 0.07)
# view-dependent formats, motion feature, weight accumulation, image bands
tr = synth.synth_tree(4, "ball"); M = tr["M"]
o, d = synth.synth_rays(900)
rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
for fmt, D in (("SH9", 28), ("SH4", 13), ("SG3", 10), ("SH4", 21)):
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format=fmt, map_location=dev)
    tree.extra_data = torch.rand(3, 4, device=dev) + 0.5
    feats = torch.randn(M, D, device=dev, requires_grad=True)
    r = sv.VolumeRenderer(tree)
    r(feats, rays, transformation_matrices=torch.eye(4, device=dev).repeat(M, 1, 1).contiguous()).sum().backward()
    r.render_persp(feats, torch.from_numpy(synth.synth_cameras(1)[0]).to(dev), width=29, height=19, fx=25.0, rows=(8, 19)).sum().backward()
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=6, map_location=dev)
feats = torch.randn(M, 6, device=dev)
r = sv.VolumeRenderer(tree)
jf = torch.randn(5, 7, device=dev, requires_grad=True)
sw = torch.rand(M, 3, device=dev); ji = torch.randint(0, 5, (M, 3), device=dev, dtype=torch.int32)
r.motion_feature_render(feats, jf, sw, ji, rays).sum().backward()
# motion feature: table form (900 rays * 32 >= M) above; staged kernels (few rays) with B = 4 and B = 3, F = 32 / 7 / 17
few = sv.Rays(*(t[:7].contiguous() for t in rays))
for F, B in ((32, 4), (7, 3), (17, 4), (32, 5)):
    jf = torch.randn(6, F, device=dev, requires_grad=True)
    sw = torch.rand(M, B, device=dev); ji = torch.randint(0, 6, (M, B), device=dev, dtype=torch.int32)
    r.motion_feature_render(feats, jf, sw, ji, few).sum().backward()
    r.motion_feature_render(feats, jf, sw, ji, rays).sum().backward()
# SG / ASG rows with three channels (lane-private RGB kernels), the point-wise operators, the dense-grid render
for fmt, D, cols in (("SG4", 13, 4), ("ASG9", 28, 11)):
    t4 = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format=fmt, map_location=dev)
    t4.extra_data = torch.rand(int(fmt[-1]), cols, device=dev) + 0.5
    fx = torch.randn(M, D, device=dev, requires_grad=True)
    sv.VolumeRenderer(t4)(fx, rays).sum().backward()
    sv.VolumeRenderer(t4)(fx, few).sum().backward()
q = torch.rand(777, 3, device=dev) * 1.2 - 0.1
fq = torch.randn(M, 6, device=dev, requires_grad=True)
tree(fq, q).sum().backward()
tree.set(q, torch.randn(777, 6, device=dev)); tree.set(q, torch.randn(777, 2, device=dev))
tree[:].corners; tree.snap(q)
cam = C.CameraSpec(); cam.c2w = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev); cam.fx = cam.fy = 31.0; cam.width, cam.height = 33, 19
C.grid_weight_render(torch.rand(17, 17, 17, device=dev) * 9 - 3, cam, r._get_options(), torch.zeros(3, device=dev), torch.ones(3, device=dev))
with tree.accumulate_weights() as acc:
    r(feats, rays); r.render_persp(feats, torch.from_numpy(synth.synth_cameras(1)[0]).to(dev), width=29, height=19, fx=25.0)
# general kernels: feature rows wider than 128 channels (float32) and the float64 instantiation
for D, dt in ((150, torch.float32), (9, torch.float64), (131, torch.float64)):
    tw = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    fw = torch.randn(M, D, device=dev, dtype=dt, requires_grad=True)
    rw = sv.Rays(*(t.to(dt) for t in rays))
    rr = sv.VolumeRenderer(tw)
    ow, dw = rr.forward_with_depth(fw, rw)
    ow.sum().backward()
    rr.render_depth(fw.detach(), rw)
    camw = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev).to(dt)
    rr.render_persp(fw, camw, width=29, height=19, fx=25.0, rows=(8, 19)).sum().backward()
    tw(fw.detach(), torch.rand(500, 3, device=dev, dtype=dt) * 1.2 - 0.1, want_node_ids=True, want_leaf_node=True)
torch.cuda.synchronize()
print("sanitize_small done, launches", C.launch_count())
