"""Config C5 (BASELINE.json): depth-10 shell octree (~30M leaves), 64-channel features, 1920x1080 views.
Checks a ray sample against the oracle and times image renders + a random-ray fwd/bwd step."""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
from oracle import oracle as orc

dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L, D = 10, 64
t0 = time.time()
tr = synth.synth_tree(L, "shell")
print(f"tree: nodes {tr['n_nodes']} leaves {tr['n_leaves']} M {tr['M']}  ({time.time()-t0:.1f} s host build)", flush=True)
M = tr["M"]
g = torch.Generator(device=dev); g.manual_seed(0)
feats = torch.randn(M, D, device=dev, generator=g)
feats[:, -1] = torch.rand(M, device=dev, generator=g) * 10 - 2
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree); opt = r._get_options()
t0 = time.time(); acc = tree.accel(feats); torch.cuda.synchronize()
print("accel:", acc.describe(), f"{time.time()-t0:.3f} s", flush=True)
ts = r._render_spec(feats, 1 << 21)      # activated table + hit marks, as VolumeRenderer attaches them

def ev(fn, warm=2, it=5):
    for _ in range(warm): fn()
    out = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); out.append(a.elapsed_time(b))
    return float(np.median(out))

cams = [torch.from_numpy(c).to(dev) for c in synth.synth_cameras(8)]
cs = [sv.renderer._make_camera_spec(c, 1920, 1080, 1500.0, 1500.0) for c in cams]
img, dep = C.volume_render_image_with_depth(ts, cs[0], opt)
print("view0 hit frac", float((img[..., -1] > 0).float().mean()), flush=True)
ms = [ev(lambda c=c: C.volume_render_image_with_depth(ts, c, opt), 1, 3) for c in cs]
print("1080p view ms:", [round(m, 2) for m in ms], "mean", round(float(np.mean(ms)), 2), "-> Mpix/s", round(2.0736 / np.mean(ms) * 1e3, 1), flush=True)
Q = 1 << 20
o, d = synth.synth_rays(Q)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
gout = torch.randn(Q, D, device=dev)
out = C.volume_render(ts, rs, opt)
print("random rays fwd ms", ev(lambda: C.volume_render(ts, rs, opt)), "bwd ms", ev(lambda: C.volume_render_backward(ts, rs, opt, gout, saved_out=out), 1, 3), flush=True)
# oracle sample
T = orc.Tree(tr["child"], tr["data"])
f_np = feats.cpu().numpy()
sel = np.arange(0, Q, Q // 1024)[:1024]
o_ref, d_ref, cnt = orc.render_rays(T, f_np, o[sel], d[sel], want_counters=True)
err = np.abs(out.cpu().numpy()[sel] - o_ref)
print("vs oracle: max", err.max(), "frac>tol", float((err > 1e-4 + 1e-3 * np.abs(o_ref)).mean()), "counters/ray", {k: cnt[k] / cnt["Q"] for k in "S LV V H".split()})
