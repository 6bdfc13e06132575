"""Timing of the view-dependent (SH9) render on the C3 tree against the reference's CUDA kernels (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
import refdrv
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q = 1 << 20; B, C = 9, 3; D = B * C + 1
tr = synth.synth_tree(8, "ball")
f = synth.synth_features(tr["M"], D); f[:, :-1] *= 0.5
o, d = synth.synth_rays(Q)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format="SH9", map_location=dev)
r = sv.VolumeRenderer(tree)
feats = torch.from_numpy(f).to(dev).requires_grad_(True)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rays = sv.Rays(o_t, d_t, d_t)
g = torch.randn(Q, C + 1, device=dev)
def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
out = r(feats, rays)
print("SH9 fwd ms", ev(lambda: r(feats.detach(), rays)))
def fb():
    feats.grad = None
    (r(feats, rays) * g).sum().backward()
print("SH9 fwd+bwd ms (autograd)", ev(fb))
if refdrv.available():
    m = refdrv.module()
    rts = refdrv.tree_spec(feats.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
    rrs = refdrv.rays_spec(o_t, d_t); ro = refdrv.options(); ro.format, ro.basis_dim, ro.min_comp, ro.max_comp = 1, B, 0, B - 1
    ref = m.volume_render(rts, rrs, ro)
    print("max |ours - ref|", float((out.detach() - ref).abs().max()))
    print("REF SH9 fwd ms", ev(lambda: m.volume_render(rts, rrs, ro), 2), "bwd ms", ev(lambda: m.volume_render_backward(rts, rrs, ro, g), 2))
