"""Generic view-dependent kernels (rows that are not the 3-channel layout): SH4 with 5 output channels, D = 21, on the C3
tree, 2^20 rays (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
Q = 1 << 20
tr = synth.synth_tree(8, "ball"); M = tr["M"]
o, d = synth.synth_rays(Q)
rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
B, Cc = 4, 5; D = B * Cc + 1
f = torch.from_numpy(synth.synth_features(M, D)).to(dev).requires_grad_(True)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format="SH4", map_location=dev)
r = sv.VolumeRenderer(tree)
g = torch.randn(Q, Cc + 1, device=dev)
def ev(fn, n=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
print("SH4 x 5 channels fwd ms", ev(lambda: r(f.detach(), rays)))
def fb():
    f.grad = None
    (r(f, rays) * g).sum().backward()
print("SH4 x 5 channels fwd+bwd ms", ev(fb))
