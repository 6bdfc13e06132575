"""Golden fixture for the view-dependent formats (SH / SG / ASG, per-row view rotation, component windows) and the
motion-feature render, produced by the UNMODIFIED reference CUDA extension on a B200:
    gpurun -- python tests/golden/make_golden_fmt.py gpurun_out/golden
Not recorded because the reference cannot produce them: motion_feature_render_backward (adds into an uninitialised
array) and NDC images (volume_render_image raises). With per-row rotations the
reference's backward keeps a stale basis in its second pass; the fixture records what it computes and the oracle
reproduces it with stale_basis=True."""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from svox_t_b200 import synth  # noqa: E402
import refdrv  # noqa: E402


def fmt_cases(M, rng):
    """name -> (format, basis_dim, C, extra, min_comp, max_comp, with_tm, sigma_thresh/stop_thresh)"""
    sg_extra = np.concatenate([rng.uniform(1.0, 5.0, (4, 1)), synth._unit(rng, 4)], 1).astype(np.float32)
    asg_extra = rng.standard_normal((2, 11)).astype(np.float32)
    asg_extra[:, :2] = np.abs(asg_extra[:, :2])
    return {
        "sh9": (1, 9, 3, None, 0, 8, False, 0.0),
        "sh9_tm": (1, 9, 3, None, 0, 8, True, 0.0),
        "sh9_window": (1, 9, 3, None, 1, 3, False, 0.0),
        "sh4_fast": (1, 4, 5, None, 0, 3, False, 1e-2),
        "sh16": (1, 16, 2, None, 0, 15, False, 0.0),
        "sh25": (1, 25, 1, None, 0, 24, False, 0.0),
        "sg4_tm": (2, 4, 3, sg_extra, 0, 3, True, 0.0),
        "asg2": (3, 2, 3, asg_extra, 0, 1, False, 0.0),
    }


if __name__ == "__main__":
    out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "_new")
    assert refdrv.available()
    dev = torch.device("cuda:0")
    m = refdrv.module()
    L, Q = 4, 640
    tr = synth.synth_tree(L, "ball")
    M = tr["M"]
    o, d = synth.synth_rays(Q, seed=7)
    rng = np.random.default_rng(21)
    vd = synth._unit(rng, Q).astype(np.float32) * rng.uniform(0.5, 1.5, (Q, 1)).astype(np.float32)   # not unit length
    tm = np.zeros((M, 4, 4), np.float32)
    for i in range(M):
        qm, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        tm[i, :3, :3] = qm
        tm[i, :3, 3] = rng.standard_normal(3)
        tm[i, 3, 3] = 1.0
    off, inv = np.zeros(3, np.float32), np.ones(3, np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    blob = dict(child=tr["child"], data=tr["data"], parent_depth=tr["parent_depth"], origins=o, dirs=d, vdirs=vd, tm=tm)
    for name, (fmt, B, C, extra, cmin, cmax, with_tm, thr) in fmt_cases(M, rng).items():
        D = C * B + 1
        f = synth.synth_features(M, D, seed=zlib.crc32(name.encode()) % 1000)
        f[:, :-1] *= 0.7
        g = rng.standard_normal((Q, C + 1)).astype(np.float32)
        ts = refdrv.tree_spec(t(f), t(tr["child"]), t(tr["data"]), t(tr["parent_depth"]), t(off), t(inv), tr["n_nodes"])
        if extra is not None:
            ts.extra_data = t(extra)
        if with_tm:
            ts.transformation_matrices = t(tm)
        rs = refdrv.rays_spec(t(o), t(d))
        rs.vdirs = t(vd)
        opt = refdrv.options(sigma_thresh=thr, stop_thresh=thr)
        opt.format, opt.basis_dim, opt.min_comp, opt.max_comp = fmt, B, cmin, cmax
        out = m.volume_render(ts, rs, opt)
        grad = m.volume_render_backward(ts, rs, opt, t(g))
        assert tuple(out.shape) == (Q, C + 1), out.shape
        blob.update({name + "_features": f, name + "_grad_out": g, name + "_ref_out": out.cpu().numpy(),
                     name + "_ref_grad": grad.cpu().numpy(),
                     name + "_meta": np.array([fmt, B, C, cmin, cmax, int(with_tm)], np.int64),
                     name + "_thresh": np.float32(thr)})
        if extra is not None:
            blob[name + "_extra"] = extra
        print(name, "ok", float(out.abs().mean()), float(grad.abs().sum()))

    # ---- motion-feature render (forward only) ----------------------------------------------------------------------
    J, F, B = 6, 12, 4
    f4 = synth.synth_features(M, 4, seed=3)
    jf = rng.standard_normal((J, F)).astype(np.float32)
    sw = rng.dirichlet(np.ones(B), M).astype(np.float32)
    sw[sw < 0.08] = 0.0
    ji = rng.integers(0, J, (M, B)).astype(np.int32)
    ts = refdrv.tree_spec(t(f4), t(tr["child"]), t(tr["data"]), t(tr["parent_depth"]), t(off), t(inv), tr["n_nodes"])
    ts.joint_features, ts.skinning_weights, ts.joint_index = t(jf), t(sw), t(ji)
    rs = refdrv.rays_spec(t(o), t(d))
    for tag, thr in (("default", 0.0), ("fast", 1e-2)):
        opt = refdrv.options(sigma_thresh=thr, stop_thresh=thr, background_brightness=0.5)
        blob["mf_ref_out_" + tag] = m.motion_feature_render(ts, rs, opt).cpu().numpy()
    blob.update(mf_features=f4, mf_jf=jf, mf_sw=sw, mf_ji=ji)
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, "y_fmt_ball_L4.npz"), **blob)
    print("fixture ok")
