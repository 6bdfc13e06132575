"""General kernels (svoxb_render_wide.cu) against the reference's CUDA kernels: float32 rows of 160 channels and the
float64 instantiation at D = 32, depth-7 ball, 2^18 random rays (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
import refdrv
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q = 1 << 18
tr = synth.synth_tree(7, "ball"); M = tr["M"]
o, d = synth.synth_rays(Q)
def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for D, dt in ((160, torch.float32), (32, torch.float64), (32, torch.float32)):
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    f = torch.from_numpy(synth.synth_features(M, D)).to(dev).to(dt)
    o_t, d_t = torch.from_numpy(o).to(dev).to(dt), torch.from_numpy(d).to(dev).to(dt)
    g = torch.randn(Q, D, device=dev, dtype=dt)
    r = sv.VolumeRenderer(tree)
    ts = r._render_spec(f, Q)
    rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t)); opt = r._get_options()
    out = C.volume_render(ts, rs, opt)
    t_f = ev(lambda: C.volume_render(ts, rs, opt))
    t_b = ev(lambda: C.volume_render_backward(ts, rs, opt, g, saved_out=out))
    line = f"D={D} {str(dt)[6:]}: here fwd {t_f:.3f} ms, bwd (incl. zeros_like) {t_b:.3f} ms"
    if refdrv.available():
        m = refdrv.module()
        off, scl = (tree.offset.to(dt), tree.invradius.to(dt))
        rts = refdrv.tree_spec(f, tree.child, tree.data, tree.parent_depth, off, scl, tree.filled, dtype=dt)
        rrs, ro = refdrv.rays_spec(o_t, d_t, dtype=dt), refdrv.options()
        line += f" | reference CUDA fwd {ev(lambda: m.volume_render(rts, rrs, ro), 2):.3f} ms, bwd {ev(lambda: m.volume_render_backward(rts, rrs, ro, g), 2):.3f} ms"
    print(line, flush=True)
