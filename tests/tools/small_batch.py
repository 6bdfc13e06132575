"""Per-stage device times of one training step on the C3 tree for short ray batches (what a strong-scaling shard of
the 2^20-ray batch looks like on one GPU): activation, hit marks, forward, gradient zero-fill, backward. Dev tool.
    python tests/tools/small_batch.py [D] [L]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(L, "ball")
f = synth.synth_features(tr["M"], D)
QMAX = 1 << 20
o, d = synth.synth_rays(QMAX)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(f).to(dev)
O, Dr = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
G = torch.randn(QMAX, D, device=dev)
opt = sv.VolumeRenderer(tree)._get_options()
ts = tree._spec(feats)
accel = ts._accel
lib = C.load_library()
ev = lambda: torch.cuda.Event(enable_timing=True)
grad = torch.zeros_like(feats)
print(f"C3 tree L={L} D={D} M={tr['M']}; per-stage ms (median of 7)")
print("    rays  tables  -    fwd    -      bwd    total   Mrays/s   vs 2^20 per-ray rate")
base = None
for Q in (1 << 20, 1 << 19, 1 << 18, 1 << 17, 1 << 16):
    o_t, d_t, g_t = O[:Q].contiguous(), Dr[:Q].contiguous(), G[:Q].contiguous()
    rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
    rows, evs = [], []
    for it in range(12):                 # queued back to back: the host runs ahead, the windows hold GPU time only
        e = [ev() for _ in range(6)]
        e[0].record()
        ts._act = C.Activated(feats, accel=accel, zero_table=grad)      # one pass: activation + hit marks + grad zero-fill
        e[1].record()
        e[2].record()
        out = C.volume_render(ts, rs, opt)
        e[3].record()
        e[4].record()
        C._check(lib.svoxb_render_rays_bwd_cost(C.ctypes.byref(ts._c()), C._ptr(o_t), C._ptr(d_t), C._ptr(d_t), Q,
                                                C.ctypes.byref(opt._c(sigma_thresh=0.0, stop_thresh=-1.0)),
                                                C._ptr(g_t), C._ptr(out), C._ptr(grad), C._ptr(rs._cost), C._stream()))
        e[5].record()
        evs.append(e)
    torch.cuda.synchronize()
    for e in evs[4:]:
        rows.append([e[i].elapsed_time(e[i + 1]) for i in range(5)] + [e[0].elapsed_time(e[5])])
    m = np.median(np.array(rows), axis=0)
    rate = Q / (m[5] * 1e-3) / 1e6
    if base is None:
        base = (m[2] + m[4]) / Q
    eff = base * Q / (m[2] + m[4])
    extra = ""
    print(f"{Q:8d}  {m[0]:.3f} {m[1]:.3f}  {m[2]:.3f}  {m[3]:.3f}  {m[4]:.3f}  {m[5]:.3f}   {rate:7.1f}   march {eff:.2f}{extra}")
