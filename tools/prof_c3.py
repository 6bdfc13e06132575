"""Small fixed workload for ncu: C3 tree (L=8 ball, D=32), Q random rays, fwd + bwd via the C-ABI mirror."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
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
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(f).to(dev)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
opt = sv.VolumeRenderer(tree)._get_options()
ts = tree._spec(feats)
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
