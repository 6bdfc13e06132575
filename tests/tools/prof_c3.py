"""Small fixed workload for ncu: C3 tree (L=8 ball, D=32), Q random rays, fwd + bwd via the C-ABI mirror."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
D = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(L, "ball")
f = synth.synth_features(tr["M"], D)
o, d = synth.synth_rays(Q)
SORT = os.environ.get("SORT", "")
if SORT:
    # host-side experiment: order rays by (direction bin, Morton of the cube entry point)
    nb = int(SORT)
    ax = np.abs(d).argmax(1); sg = (d[np.arange(Q), ax] > 0).astype(np.int64)
    uv = np.stack([d[np.arange(Q), (ax + 1) % 3], d[np.arange(Q), (ax + 2) % 3]], 1) / np.abs(d[np.arange(Q), ax])[:, None]
    cell = np.clip(((uv + 1) * 0.5 * nb).astype(np.int64), 0, nb - 1)
    dbin = ((ax * 2 + sg) * nb + cell[:, 0]) * nb + cell[:, 1]
    inv = 1.0 / (d.astype(np.float64) + 1e-9)
    t1 = -o * inv; t2 = t1 + inv
    tmin = np.maximum(0, np.minimum(t1, t2).max(1))
    pin = np.clip(o + tmin[:, None] * d, 0, 1 - 1e-6)
    I = (pin * 1024).astype(np.int64)
    def part(x):
        x = x & 0x3ff; x = (x | (x << 16)) & 0x30000ff; x = (x | (x << 8)) & 0x300f00f
        x = (x | (x << 4)) & 0x30c30c3; x = (x | (x << 2)) & 0x9249249; return x
    mort = (part(I[:, 0]) << 2) | (part(I[:, 1]) << 1) | part(I[:, 2])
    order = np.argsort((dbin << 30) | mort, kind="stable")
    o, d = o[order].copy(), d[order].copy()
    print("sorted into", 6 * nb * nb, "direction bins")
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(f).to(dev)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
opt = sv.VolumeRenderer(tree)._get_options()
ts = tree._spec(feats)
if os.environ.get("ACT", "1") == "1":
    ts._act = tree.activated(feats)          # what VolumeRenderer attaches for large batches
    if os.environ.get("MARKS", "1") == "1":
        ts._accel.mark_hits(feats)
g = torch.randn(Q, D, device=dev)
for _ in range(iters):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = C.volume_render(ts, rs, opt)
    e1.record()
    grad = C.volume_render_backward(ts, rs, opt, g, saved_out=out)
    e2.record()
    torch.cuda.synchronize()
    print(f"Q={Q} fwd {e0.elapsed_time(e1):.3f} ms  bwd(+zero) {e1.elapsed_time(e2):.3f} ms")
