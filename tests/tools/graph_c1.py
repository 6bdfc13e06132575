"""CUDA-graph capture of a small training step (config C1: depth-4 tree, D=16, 4096 rays): the eager step is host-bound
(ctypes + autograd, ~170 us), the captured one replays the same kernels without the host in the loop."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L, shape, D, Q = (int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (4, "all", 16, 4096)
tr = synth.synth_tree(L, shape)
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
    return out

s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        out_eager = step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
grad_eager = feats.grad.clone(); out_eager = out_eager.detach().clone()

graph = torch.cuda.CUDAGraph()
feats.grad = None
with torch.cuda.graph(graph):
    out_g = r(feats, rays)
    out_g.backward(g)
grad_static = feats.grad
for _ in range(3): graph.replay()
torch.cuda.synchronize()
print("graph == eager: out", bool(torch.equal(out_g.detach(), out_eager)),
      " grad rel", float((grad_static - grad_eager).norm() / grad_eager.norm()))
# new inputs through the static buffers
o2, d2 = synth.synth_rays(Q, seed=9)
rays.origins.copy_(torch.from_numpy(o2)); rays.dirs.copy_(torch.from_numpy(d2))
graph.replay(); torch.cuda.synchronize()
ref = step(); torch.cuda.synchronize()
print("after input update: out equal", bool(torch.equal(out_g.detach(), ref.detach())))

def timeit(fn, n=300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6
print(f"eager step {timeit(step):.1f} us, graph replay {timeit(graph.replay):.1f} us  (Q={Q}, D={D}, L={L})")
