"""Config C1 (depth-4 full tree, D=16, 4096 rays): wall-clock per fwd+bwd step through the public API, synchronised
every step -- the host-side cost of a small call (ctypes marshalling, cache look-ups, autograd) next to the kernels."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(4, "all")
D, Q = 16, 4096
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(synth.synth_features(tr["M"], D)).to(dev).requires_grad_(True)
o, d = synth.synth_rays(Q)
rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
r = sv.VolumeRenderer(tree)
g = torch.randn(Q, D, device=dev)

def step():
    feats.grad = None
    out = r(feats, rays)
    out.backward(g)

for _ in range(20): step()
torch.cuda.synchronize()
for n in (200,):
    t0 = time.perf_counter()
    for _ in range(n):
        step()
        torch.cuda.synchronize()
    print(f"C1 fwd+bwd: {(time.perf_counter() - t0) / n * 1e6:.1f} us per step (wall, synchronised)")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): step()
    b.record(); torch.cuda.synchronize()
    print(f"C1 fwd+bwd: {a.elapsed_time(b) / n * 1e3:.1f} us per step (device timeline, back to back)")
if len(sys.argv) > 1:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): step()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
