"""Golden fixture for the march variants (opacity_render, motion_render), produced by the UNMODIFIED reference CUDA
extension on a B200:   gpurun -- python tests/golden/make_golden_x.py gpurun_out/golden
The reference's opacity_render_backward launches the wrong kernel, so no reference output exists
for it; the oracle restates the code the reference wrote for it (rt_kernel.cu:562-651) and is checked by finite
differences instead."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from svox_t_b200 import synth  # noqa: E402
import refdrv  # noqa: E402

if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_new")
    assert refdrv.available()
    dev = torch.device("cuda:0")
    m = refdrv.module()
    L, D, Q, J = 5, 8, 768, 6
    tr = synth.synth_tree(L, "ball")
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q, seed=3)
    extra = np.random.default_rng(9).random((J, 3)).astype(np.float32)
    radius, center = 0.8, 0.4
    inv = np.full(3, 0.5 / radius, np.float32)
    off = np.full(3, 0.5 * (1.0 - center / radius), np.float32)
    o = ((o - 0.5) * (2 * radius) + center).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ts = refdrv.tree_spec(t(f), t(tr["child"]), t(tr["data"]), t(tr["parent_depth"]), t(off), t(inv), tr["n_nodes"])
    ts.extra_data = t(extra)
    rs = refdrv.rays_spec(t(o), t(d))
    res = {}
    for tag, (st, sp) in {"default": (0.0, 0.0), "fast": (1e-2, 1e-2)}.items():
        opt = refdrv.options(sigma_thresh=st, stop_thresh=sp)
        res["opacity_" + tag] = m.opacity_render(ts, rs, opt).cpu().numpy()[:, 0]
        mo = m.motion_render(ts, rs, opt)
        res["motion_out_" + tag] = mo[0].cpu().numpy()
        res["motion_depth_" + tag] = mo[1].cpu().numpy()[:, 0]
        res["motion_hit_" + tag] = mo[2].cpu().numpy()
        res["motion_idx_" + tag] = mo[3].cpu().numpy()[:, 0]
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "x_ball_L5_D8_variants.npz"), child=tr["child"], data=tr["data"],
                        parent_depth=tr["parent_depth"], features=f, offset=off, scaling=inv, origins=o, dirs=d,
                        extra=extra, **res)
    print("ok", {k: v.shape for k, v in res.items()})

    # ---- point kernels: LBS warp + p2v splat, forward and backward, from the reference extension ----------------------
    rng = np.random.default_rng(12)
    P, Jn, B, nv, kr, cr = 3000, 6, 4, 24, 1.5 / 24, 2.0 / 24
    Tm, w, ji = synth.synth_skeleton(P, J=Jn, B=B, seed=5)
    w[::5, 2] = 0.0
    pts = (0.5 + 0.5 * (rng.random((P, 3)) - 0.5)).astype(np.float32)
    feat = rng.random((P, 3)).astype(np.float32) * 4 - 1
    corner, size = np.zeros(3, np.float32), np.ones(3, np.float32)
    g_c = rng.standard_normal((P, 3)).astype(np.float32)
    g_m = rng.standard_normal((P, 4, 4)).astype(np.float32)
    g_v = rng.standard_normal((nv, nv, nv, 1)).astype(np.float32)
    co, mats = m.warp_vertices(t(Tm), t(pts), t(w), t(ji))
    gx, gT, gw = m.warp_vertices_backward(t(Tm), t(pts), t(w), t(ji), t(g_c), t(g_m))
    vox = m.p2v(co, t(feat), t(corner), t(size), nv, kr, cr)
    gp, gf = m.p2v_backward(t(g_v), co, t(feat), t(corner), t(size), nv, kr, cr)
    np.savez_compressed(os.path.join(out_dir, "x_points_lbs_p2v.npz"), T=Tm, pts=pts, w=w, ji=ji, feat=feat,
                        corner=corner, size=size, n_voxels=nv, kernel_radius=np.float32(kr), conv_radius=np.float32(cr),
                        g_coords=g_c, g_mats=g_m, g_vox=g_v,
                        ref_coords=co.cpu().numpy(), ref_mats=mats.cpu().numpy(), ref_gx=gx.cpu().numpy(),
                        ref_gT=gT.cpu().numpy(), ref_gw=gw.cpu().numpy(), ref_vox=vox.cpu().numpy(),
                        ref_gp=gp.cpu().numpy(), ref_gf=gf.cpu().numpy())
    print("points fixture ok")
