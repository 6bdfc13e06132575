"""Sigma-only marches (render_depth / opacity_render fwd + bwd / motion_render) on the C3 tree with D = 32 features, 2^20
random rays: plain (sigma gathered from the [M, D] table) against compact sigma array + hit marks, and the reference's
CUDA kernels. Dev tool."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
import refdrv
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q, D = 1 << 20, int(sys.argv[1]) if len(sys.argv) > 1 else 32
tr = synth.synth_tree(8, "ball"); M = tr["M"]
o, d = synth.synth_rays(Q)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rays = sv.Rays(o_t, d_t, d_t)
rs = sv.renderer._rays_spec_from_rays(rays)
f = torch.from_numpy(synth.synth_features(M, D)).to(dev)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
tree.extra_data = torch.rand(24, 3, device=dev)
r = sv.VolumeRenderer(tree); opt = r._get_options()
g = torch.randn(Q, 1, device=dev)
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
plain = tree._spec(f)
acc = plain._accel
acc._marks_key = None                      # plain: no marks, no compact sigma
fast = r._sigma_spec(f, Q)                 # compact sigma + marks
for name, ts in (("plain", plain), ("compact sigma + marks", fast)):
    if name == "plain":
        acc._marks_key = None
    else:
        acc.mark_hits(f)
    t_op = ev(lambda: C.opacity_render(ts, rs, opt))
    t_ob = ev(lambda: C.opacity_render_backward(ts, rs, opt, g))
    so = C.opacity_render(ts, rs, opt)
    t_ob1 = ev(lambda: C.opacity_render_backward(ts, rs, opt, g, saved_out=so))
    g2, g1 = C.opacity_render_backward(ts, rs, opt, g), C.opacity_render_backward(ts, rs, opt, g, saved_out=so)
    print(f"      backward with the saved forward output {t_ob1:.3f} ms; rel diff vs two-pass {float((g1 - g2).norm() / g2.norm()):.2e}")
    t_d = ev(lambda: C.render_depth(ts, rs, opt))
    t_m = ev(lambda: C.motion_render(ts, rs, opt))
    print(f"D={D} {name:24s}: opacity fwd {t_op:.3f} ms  bwd (incl. zeros_like) {t_ob:.3f} ms  depth {t_d:.3f} ms  motion {t_m:.3f} ms", flush=True)
a = C.opacity_render(plain, rs, opt); acc.mark_hits(f); b = C.opacity_render(fast, rs, opt)
print("opacity identical:", bool(torch.equal(a, b)), " depth identical:", bool(torch.equal(C.render_depth(plain, rs, opt), C.render_depth(fast, rs, opt))))
if refdrv.available():
    m = refdrv.module()
    rts = refdrv.tree_spec(f, tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
    rts.extra_data = tree.extra_data
    rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
    print(f"D={D} reference CUDA           : opacity fwd {ev(lambda: m.opacity_render(rts, rrs, ro), 3):.3f} ms  depth {ev(lambda: m.render_depth(rts, rrs, ro), 3):.3f} ms  "
          f"motion {ev(lambda: m.motion_render(rts, rrs, ro), 3):.3f} ms")
    print("opacity max |ours - ref|", float((b - m.opacity_render(rts, rrs, ro)).abs().max()))
