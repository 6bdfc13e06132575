"""Odd feature widths (scalar-lane kernels) against the reference's CUDA kernels, C3 tree (dev tool)."""
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
for D in (5, 17, 33, 65):
    f = torch.from_numpy(synth.synth_features(M, D)).to(dev).requires_grad_(True)
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    r = sv.VolumeRenderer(tree)
    g = torch.randn(Q, D, device=dev)
    fw = ev(lambda: r(f.detach(), rays))
    def fb():
        f.grad = None
        (r(f, rays) * g).sum().backward()
    line = f"D={D}: fwd {fw:.2f} ms, fwd+bwd {ev(fb):.2f} ms"
    if m is not None:
        rts = refdrv.tree_spec(f.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
        rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
        line += f" | REF fwd {ev(lambda: m.volume_render(rts, rrs, ro), 2):.2f} ms, bwd {ev(lambda: m.volume_render_backward(rts, rrs, ro, g), 2):.2f} ms"
    print(line, flush=True)
