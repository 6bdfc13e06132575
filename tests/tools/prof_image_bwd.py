"""Camera-ray forward + backward on the C3 tree (D = 32): 800x800 (config C2's view) and 1920x1080, through
VolumeRenderer.render_persp + autograd. Prints CUDA-event times (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
D = 32
tr = synth.synth_tree(8, "ball")
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(synth.synth_features(tr["M"], D)).to(dev).requires_grad_(True)
r = sv.VolumeRenderer(tree)
cam = torch.from_numpy(synth.synth_cameras(1, dist=1.0)[0]).to(dev)
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for W, H, fx in ((800, 800, 1111.111), (1920, 1080, 1500.0)):
    g = torch.randn(H, W, D, device=dev)
    fwd = ev(lambda: r.render_persp(feats.detach(), cam, width=W, height=H, fx=fx))
    def fb():
        feats.grad = None
        (r.render_persp(feats, cam, width=W, height=H, fx=fx) * g).sum().backward()
    print(f"{W}x{H}: fwd {fwd:.3f} ms, fwd+bwd (autograd, incl. the loss product) {ev(fb):.3f} ms", flush=True)
