"""Fixed workload for ncu on config C5's scene: depth-10 shell octree (32.9 M rows x 64 channels, 8.4 GB of features,
three-stage accelerator), 2^20 random rays fwd + bwd and one 1920x1080 view -- the D = 64 quad kernels
(march_*_quad_kernel<16, 1, ...>). Prints CUDA-event times; under ncu use -k regex:march -c N.
    python tests/tools/prof_c5.py [iters] [Q] [thread caps, e.g. 512,576,704,768: sweep of the CTA size (dev knobs)]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L, D = 10, 64
t0 = time.time()
tr = synth.synth_tree(L, "shell")
M = tr["M"]
g = torch.Generator(device=dev).manual_seed(0)
feats = torch.randn(M, D, device=dev, generator=g)
feats[:, -1] = torch.rand(M, device=dev, generator=g) * 10 - 2
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree); opt = r._get_options()
ts = r._render_spec(feats, 1 << 21)
print(f"scene ready in {time.time() - t0:.1f} s: M={M}, accel {ts._accel.describe()}", flush=True)
o, d = synth.synth_rays(Q)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
gout = torch.randn(Q, D, device=dev)
grad = torch.zeros_like(feats)
lib, bopt = C.load_library(), opt._c(sigma_thresh=0.0, stop_thresh=-1.0)
cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
cs = sv.renderer._make_camera_spec(cam, 1920, 1080, 1500.0, 1500.0)
ev = lambda: torch.cuda.Event(enable_timing=True)
def run(tag=""):
    for _ in range(iters):
        e = [ev() for _ in range(4)]
        e[0].record()
        out = C.volume_render(ts, rs, opt)
        e[1].record()
        C._check(lib.svoxb_render_rays_bwd_cost(C.ctypes.byref(ts._c()), C._ptr(o_t), C._ptr(d_t), C._ptr(d_t), Q,
                                                C.ctypes.byref(bopt), C._ptr(gout), C._ptr(out), C._ptr(grad),
                                                C._ptr(rs._cost), C._stream()))
        e[2].record()
        C.volume_render_image_with_depth(ts, cs, opt)
        e[3].record()
        torch.cuda.synchronize()
        print(f"C5 Q={Q}{tag}: fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms  "
              f"1080p view+depth {e[2].elapsed_time(e[3]):.3f} ms", flush=True)


run()
for cap in (sys.argv[3].split(",") if len(sys.argv) > 3 else []):
    os.environ["SVOXB_FWD_THREADS_CAP"] = os.environ["SVOXB_BWD_THREADS_CAP"] = cap
    run(f" threads<={cap}")
