"""Timing of the render variants against the reference's CUDA kernels on the C3 tree (dev tool): motion-feature render,
opacity render, depth, motion render, SG format, point query."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
import refdrv
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q = 1 << 20
tr = synth.synth_tree(8, "ball"); M = tr["M"]
o, d = synth.synth_rays(Q)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rays = sv.Rays(o_t, d_t, d_t)
m = refdrv.module() if refdrv.available() else None
def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
rng = np.random.default_rng(0)
# ---- motion feature -------------------------------------------------------------------------------------------------
D = 4; J, F, B = 24, 32, 4
f = torch.from_numpy(synth.synth_features(M, D)).to(dev)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
tree.extra_data = torch.rand(J, 3, device=dev)
r = sv.VolumeRenderer(tree)
jf = torch.randn(J, F, device=dev, requires_grad=True)
sw = torch.from_numpy(rng.dirichlet(np.ones(B), M).astype(np.float32)).to(dev)
ji = torch.from_numpy(rng.integers(0, J, (M, B)).astype(np.int32)).to(dev)
g = torch.randn(Q, F, device=dev)
out = r.motion_feature_render(f, jf, sw, ji, rays)
print("motion_feature fwd ms", ev(lambda: r.motion_feature_render(f, jf.detach(), sw, ji, rays)))
def fb():
    jf.grad = None
    (r.motion_feature_render(f, jf, sw, ji, rays) * g).sum().backward()
print("motion_feature fwd+bwd ms", ev(fb))
print("opacity fwd ms", ev(lambda: r.opacity_render(f, rays)), " depth ms", ev(lambda: r.render_depth(f, rays)),
      " motion_render ms", ev(lambda: r.motion_render(f, rays)))
pts = torch.rand(Q, 3, device=dev)
print("query 1M pts ms", ev(lambda: tree(f, pts, want_node_ids=True, want_data_ids=True)))
if m is not None:
    rts = refdrv.tree_spec(f, tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
    rts.joint_features, rts.skinning_weights, rts.joint_index = jf.detach(), sw, ji
    rts.extra_data = tree.extra_data
    rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
    ref = m.motion_feature_render(rts, rrs, ro)
    print("  max |ours - ref|", float((out.detach() - ref).abs().max()))
    print("  REF motion_feature fwd ms", ev(lambda: m.motion_feature_render(rts, rrs, ro), 2),
          "bwd ms", ev(lambda: m.motion_feature_render_backward(rts, rrs, ro, g), 2))
    print("  REF opacity fwd ms", ev(lambda: m.opacity_render(rts, rrs, ro), 2), " depth ms", ev(lambda: m.render_depth(rts, rrs, ro), 2),
          " motion_render ms", ev(lambda: m.motion_render(rts, rrs, ro), 2))
    print("  REF query 1M pts ms", ev(lambda: m.query_vertical(rts, pts), 2))
# ---- SG format (generic kernels) ------------------------------------------------------------------------------------
Bs, Cs = 4, 3; D = Bs * Cs + 1
fs = torch.from_numpy(synth.synth_features(M, D)).to(dev).requires_grad_(True)
tree2 = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format="SG4", map_location=dev)
extra = np.concatenate([rng.uniform(1, 5, (Bs, 1)), synth._unit(rng, Bs)], 1).astype(np.float32)
tree2.extra_data = torch.from_numpy(extra).to(dev)
r2 = sv.VolumeRenderer(tree2)
gs = torch.randn(Q, Cs + 1, device=dev)
print("SG4 fwd ms", ev(lambda: r2(fs.detach(), rays)))
def fb2():
    fs.grad = None
    (r2(fs, rays) * gs).sum().backward()
print("SG4 fwd+bwd ms", ev(fb2))
if m is not None:
    rts = refdrv.tree_spec(fs.detach(), tree2.child, tree2.data, tree2.parent_depth, tree2.offset, tree2.invradius, tree2.filled)
    rts.extra_data = tree2.extra_data
    ro = refdrv.options(); ro.format, ro.basis_dim, ro.min_comp, ro.max_comp = 2, Bs, 0, Bs - 1
    print("  REF SG4 fwd ms", ev(lambda: m.volume_render(rts, rrs, ro), 2), "bwd ms", ev(lambda: m.volume_render_backward(rts, rrs, ro, gs), 2))
