"""Motion-feature render on the C3 tree for ncu / timing: Q rays, J=24, F=32, B=4 (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tr = synth.synth_tree(8, "ball"); M = tr["M"]
o, d = synth.synth_rays(Q)
rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
rng = np.random.default_rng(0)
D, J, F, B = 4, 24, 32, 4
f = torch.from_numpy(synth.synth_features(M, D)).to(dev)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree)
jf = torch.randn(J, F, device=dev, requires_grad=True)
sw = torch.from_numpy(rng.dirichlet(np.ones(B), M).astype(np.float32)).to(dev)
ji = torch.from_numpy(rng.integers(0, J, (M, B)).astype(np.int32)).to(dev)
g = torch.randn(Q, F, device=dev)
for _ in range(iters):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    jf.grad = None
    e[0].record()
    out = r.motion_feature_render(f, jf, sw, ji, rays)
    e[1].record()
    (out * g).sum().backward()
    e[2].record()
    torch.cuda.synchronize()
    print(f"Q={Q} mf fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd(+torch mul/sum) {e[1].elapsed_time(e[2]):.3f} ms")
