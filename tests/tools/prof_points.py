"""Timing of the point kernels and their backwards (P = 2^20 points, J = 24, B = 4; 256^3 splat)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
import refdrv
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
P = 1 << 20
rng = np.random.default_rng(2)
vox = synth._occupied_keys(8, "ball")
pts = synth.voxel_centers(vox[rng.permutation(len(vox))[:P]], 8)
Tm, w, ji = synth.synth_skeleton(P)
p, Tm_t, w_t, ji_t = (torch.from_numpy(a).to(dev) for a in (pts, Tm, w, ji))
f = torch.from_numpy(synth.synth_features(P, 8)).to(dev)
corner, size = torch.zeros(3, device=dev), torch.ones(3, device=dev)
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
gx, gm = torch.randn(P, 3, device=dev), torch.randn(P, 4, 4, device=dev)
print("warp_vertices ms", ev(lambda: C.warp_vertices(Tm_t, p, w_t, ji_t)))
print("warp_vertices_backward ms", ev(lambda: C.warp_vertices_backward(Tm_t, p, w_t, ji_t, gx, gm)))
warped = C.warp_vertices(Tm_t, p, w_t, ji_t)[0]
print("p2v ms", ev(lambda: C.p2v(warped, f, corner, size, 256, 1.5 / 256, 2.0 / 256)))
gv = torch.randn(256, 256, 256, 1, device=dev)
print("p2v_backward ms", ev(lambda: C.p2v_backward(gv, warped, f, corner, size, 256, 1.5 / 256, 2.0 / 256)))
if refdrv.available():
    m = refdrv.module()
    print("REF warp_vertices_backward ms", ev(lambda: m.warp_vertices_backward(Tm_t, p, w_t, ji_t, gx, gm), 3))
    print("REF p2v_backward ms", ev(lambda: m.p2v_backward(gv, warped, f, corner, size, 256, 1.5 / 256, 2.0 / 256), 3))
