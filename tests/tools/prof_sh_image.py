"""SH9 (PlenOctree layout, D = 28) 1920x1080 views of the C3 ball and of a depth-9 shell: the three-channel
view-dependent image kernel (sh_rgb_fwd_kernel<9, ..., IMAGE>). Prints CUDA-event times (dev tool)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
W, H, fx = 1920, 1080, 1500.0
def ev(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for L, shape, fmt, B, C in ((8, "ball", "SH9", 9, 3), (9, "shell", "SH9", 9, 3), (8, "ball", "SH4", 4, 5)):
    D = B * C + 1          # C = 3: the PlenOctree layout (sh_rgb kernels); any other C: the general view-dependent kernels
    tr = synth.synth_tree(L, shape)
    f = synth.synth_features(tr["M"], D); f[:, :-1] *= 0.5
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, data_format=fmt, map_location=dev)
    r = sv.VolumeRenderer(tree)
    feats = torch.from_numpy(f).to(dev)
    ms = []
    for c2w in synth.synth_cameras(4, dist=1.0):
        cam = torch.from_numpy(c2w).to(dev)
        ms.append(ev(lambda: r.render_persp(feats, cam, width=W, height=H, fx=fx)))
    print(f"{fmt} x {C} channels 1080p view, depth-{L} {shape} ({tr['M']} rows): {np.mean(ms):.3f} ms (4 views: {', '.join('%.3f' % m for m in ms)})", flush=True)
