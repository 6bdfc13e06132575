"""Dev check on a B200: view-dependent formats + motion-feature render, CUDA path vs CPU oracle (and the golden fixture
if present). Run: gpurun -- python tools/check_fmt.py [fixture.npz]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import csrc as C
from oracle import oracle as orc

dev = torch.device("cuda:0"); torch.cuda.set_device(0)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "golden", "y_fmt_ball_L4.npz")
z = np.load(path)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
T = orc.Tree(z["child"], z["data"])
names = sorted(k[:-5] for k in z.files if k.endswith("_meta"))
NAMES = {1: "SH", 2: "SG", 3: "ASG"}
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
for name in names:
    fmt, B, Cc, cmin, cmax, with_tm = (int(v) for v in z[name + "_meta"])
    f, g = z[name + "_features"], z[name + "_grad_out"]
    extra = z[name + "_extra"] if name + "_extra" in z.files else None
    tm = z["tm"] if with_tm else None
    thr = float(z[name + "_thresh"])
    for accel in (True, False):
        tree = sv.N3Tree.from_tensors(z["child"], z["data"], z["parent_depth"], data_dim=f.shape[1],
                                      data_format=f"{NAMES[fmt]}{B}", map_location=dev)
        if extra is not None: tree.extra_data = cu(extra)
        if not accel: tree.accel = lambda *a, **k: None
        r = sv.VolumeRenderer(tree, min_comp=cmin, max_comp=cmax)
        r.sigma_thresh = thr; r.stop_thresh = thr
        feats = cu(f).requires_grad_(True)
        rays = sv.Rays(cu(z["origins"]), cu(z["dirs"]), cu(z["vdirs"]))
        out = r(feats, rays, transformation_matrices=cu(tm) if with_tm else None)
        (out * cu(g)).sum().backward()
        out_n, grad_n = out.detach().cpu().numpy(), feats.grad.cpu().numpy()
        o_orc = orc.render_rays_fmt(T, f, z["origins"], z["dirs"], z["vdirs"], fmt, B, extra=extra, tm=tm, min_comp=cmin,
                                    max_comp=cmax, sigma_thresh=thr, stop_thresh=thr)
        g_orc = orc.render_rays_fmt_backward(T, f, z["origins"], z["dirs"], z["vdirs"], g, fmt, B, extra=extra, tm=tm,
                                             min_comp=cmin, max_comp=cmax)
        g_orc_stale = orc.render_rays_fmt_backward(T, f, z["origins"], z["dirs"], z["vdirs"], g, fmt, B, extra=extra, tm=tm,
                                                   min_comp=cmin, max_comp=cmax, stale_basis=True)
        print(f"{name:12s} accel={int(accel)} fwd cuda-orc {np.abs(out_n-o_orc).max():.2e} cuda-ref {np.abs(out_n-z[name+'_ref_out']).max():.2e} "
              f"orc-ref {np.abs(o_orc-z[name+'_ref_out']).max():.2e} | grad cuda-orc {rel(grad_n,g_orc):.2e} "
              f"orc(stale)-ref {rel(g_orc_stale, z[name+'_ref_grad']):.2e} cuda-ref {rel(grad_n, z[name+'_ref_grad']):.2e}")
# motion feature
f4, jf, sw, ji = z["mf_features"], z["mf_jf"], z["mf_sw"], z["mf_ji"]
tree = sv.N3Tree.from_tensors(z["child"], z["data"], z["parent_depth"], data_dim=4, map_location=dev)
rays = sv.Rays(cu(z["origins"]), cu(z["dirs"]), cu(z["dirs"]))
for tag, thr in (("default", 0.0), ("fast", 1e-2)):
    r = sv.VolumeRenderer(tree, background_brightness=0.5); r.sigma_thresh = thr; r.stop_thresh = thr
    jft = cu(jf).requires_grad_(True)
    out = r.motion_feature_render(cu(f4), jft, cu(sw), cu(ji), rays)
    g = np.random.default_rng(1).standard_normal(out.shape).astype(np.float32)
    (out * cu(g)).sum().backward()
    o_orc = orc.motion_feature_render(T, f4, z["origins"], z["dirs"], jf, sw, ji, background_brightness=0.5, sigma_thresh=thr, stop_thresh=thr)
    g_orc = orc.motion_feature_render_backward(T, f4, z["origins"], z["dirs"], jf, sw, ji, g)
    print(f"mf {tag}: fwd cuda-orc {np.abs(out.detach().cpu().numpy()-o_orc).max():.2e} cuda-ref {np.abs(out.detach().cpu().numpy()-z['mf_ref_out_'+tag]).max():.2e} "
          f"grad cuda-orc {rel(jft.grad.cpu().numpy(), g_orc):.2e}")
