"""Development check on a B200: parity of the CUDA path against the C oracle and the compiled reference,
plus first timings. Run: gpurun -- python tools/gpu_check.py"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
from oracle import oracle as orc
import refdrv

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), "ref available:", refdrv.available())

def stats(name, a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    tol = 1e-4 + 1e-3 * np.abs(b)
    print(f"  {name}: max abs {err.max():.3e} mean abs {err.mean():.3e} frac>tol {(err > tol).mean():.3e} "
          f"relL2 {np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30):.3e}")

def ev_time(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))

def scene(L, shape, D, Q, use_accel=True):
    tr = synth.synth_tree(L, shape)
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q)
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    feats = torch.from_numpy(f).to(dev).requires_grad_(True)
    return tr, f, o, d, tree, feats

def check(L, shape, D, Q, label):
    print(f"== {label}: L={L} {shape} D={D} Q={Q}")
    tr, f, o, d, tree, feats = scene(L, shape, D, Q)
    T = orc.Tree(tr["child"], tr["data"])
    o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    rays = sv.Rays(o_t, d_t, d_t)
    r = sv.VolumeRenderer(tree)
    acc = tree.accel(feats)
    print("  accel:", acc.describe() if acc is not None else None)
    out, depth = r.forward_with_depth(feats, rays)
    g = np.random.default_rng(5).standard_normal(out.shape).astype(np.float32)
    g_t = torch.from_numpy(g).to(dev)
    (out * g_t).sum().backward()
    grad = feats.grad.detach().cpu().numpy()
    out_n, depth_n = out.detach().cpu().numpy(), depth.detach().cpu().numpy()[:, 0]
    # generic (no accel) path
    ts = tree._spec(feats, _with_accel=False)
    out_g, depth_g = C._render_fwd(ts, sv.renderer._rays_spec_from_rays(rays), r._get_options(), True)
    grad_g = C.volume_render_backward(ts, sv.renderer._rays_spec_from_rays(rays), r._get_options(), g_t, saved_out=out_g)
    print("  accel vs generic: out maxdiff", float((out_g - out).abs().max()), "depth", float((depth_g - depth).abs().max()),
          "grad relL2", float((grad_g - feats.grad).norm() / feats.grad.norm()))
    n_or = min(Q, 8192)
    oo, od = orc.render_rays(T, f, o[:n_or], d[:n_or])[:2]
    stats("fwd vs oracle32", out_n[:n_or], oo); stats("depth vs oracle32", depth_n[:n_or], od)
    if Q <= 8192:
        og = orc.render_rays_backward(T, f, o, d, g)
        stats("grad vs oracle32", grad, og)
    o64 = orc.render_rays(T, f, o[:n_or], d[:n_or], dtype=np.float64)[0]
    stats("fwd vs oracle64", out_n[:n_or], o64); stats("oracle32 vs oracle64", oo, o64)
    if refdrv.available():
        m = refdrv.module()
        rts = refdrv.tree_spec(feats.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
        rrs = refdrv.rays_spec(o_t, d_t); ro = refdrv.options()
        rout = m.volume_render(rts, rrs, ro); rdepth = m.render_depth(rts, rrs, ro)
        rgrad = m.volume_render_backward(rts, rrs, ro, g_t)
        stats("fwd vs REF cuda", out_n, rout.cpu().numpy()); stats("depth vs REF cuda", depth_n, rdepth.cpu().numpy()[:, 0])
        stats("grad vs REF cuda", grad, rgrad.cpu().numpy())
        stats("oracle32 vs REF cuda (fwd)", oo, rout.cpu().numpy()[:n_or])
        # query parity
        pts = torch.rand(20000, 3, device=dev)
        v, nid, did, leaf = tree.forward(feats.detach(), pts, want_node_ids=True, want_data_ids=True, want_leaf_node=True)
        rv, rnid, rdid, rleaf = m.query_vertical(rts, pts)
        valid = did >= 0
        print("  query: node_ids equal", bool((nid == rnid).all()), "data_ids equal(valid)", bool((did[valid] == rdid[valid]).all()),
              "values equal(valid)", bool((v[valid] == rv[valid]).all()),
              "leafset equal", bool(torch.equal(leaf, rleaf[torch.argsort(sv.N3Tree._pack_index(tree, rleaf))])) if rleaf.numel() else None)
        return dict(rts=rts, rrs=rrs, ro=ro, m=m, g_t=g_t)
    return {}

check(4, "all", 16, 4096, "C1")
check(6, "ball", 33, 4096, "odd-D")
check(5, "ball", 64, 2048, "D64")

print("== timing C3: L=8 ball D=32, 1M rays")
Q = 1 << 20
tr, f, o, d, tree, feats = scene(8, "ball", 32, Q)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
rays = sv.Rays(o_t, d_t, d_t); rs = sv.renderer._rays_spec_from_rays(rays)
r = sv.VolumeRenderer(tree); opt = r._get_options()
t0 = time.time(); acc = tree.accel(feats); torch.cuda.synchronize(); print("  accel build s:", time.time() - t0, acc.describe())
ts = tree._spec(feats.detach())
g_t = torch.randn(Q, 32, device=dev)
out = C.volume_render(ts, rs, opt)
print("  fwd  (accel)  ms best/med:", ev_time(lambda: C.volume_render(ts, rs, opt)))
print("  fwd+depth     ms best/med:", ev_time(lambda: C.volume_render_with_depth(ts, rs, opt)))
print("  bwd  (accel)  ms best/med:", ev_time(lambda: C.volume_render_backward(ts, rs, opt, g_t, saved_out=out)))
tsg = tree._spec(feats.detach(), _with_accel=False)
print("  fwd  (generic) ms:", ev_time(lambda: C.volume_render(tsg, rs, opt), 1, 3))
print("  bwd  (generic) ms:", ev_time(lambda: C.volume_render_backward(tsg, rs, opt, g_t, saved_out=out), 1, 3))
print("  depth kernel ms:", ev_time(lambda: C.render_depth(ts, rs, opt)))
if refdrv.available():
    m = refdrv.module()
    rts = refdrv.tree_spec(feats.detach(), tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
    rrs = refdrv.rays_spec(o_t, d_t); ro = refdrv.options()
    print("  REF fwd ms:", ev_time(lambda: m.volume_render(rts, rrs, ro), 1, 3))
    print("  REF bwd ms:", ev_time(lambda: m.volume_render_backward(rts, rrs, ro, g_t), 1, 3))
    print("  REF depth ms:", ev_time(lambda: m.render_depth(rts, rrs, ro), 1, 3))
    rout = m.volume_render(rts, rrs, ro)
    stats("C3 fwd vs REF cuda", out.cpu().numpy(), rout.cpu().numpy())
    gr = C.volume_render_backward(ts, rs, opt, g_t, saved_out=out); rgr = m.volume_render_backward(rts, rrs, ro, g_t)
    stats("C3 grad vs REF cuda", gr.cpu().numpy(), rgr.cpu().numpy())
print("== image C2: 800x800")
cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
img, dep = r.render_persp_with_depth(feats.detach(), cam)
print("  image", tuple(img.shape), "opacity mean", float(img[..., -1].mean()), "hit frac", float((img[..., -1] > 0).float().mean()))
cs = sv.renderer._make_camera_spec(cam, 800, 800, 1111.111, 1111.111)
print("  image fwd ms:", ev_time(lambda: C.volume_render_image(ts, cs, opt)))
oc, dc = orc.camera_rays(cam.cpu().numpy(), 1111.111, 1111.111, 800, 800)
rays_c = sv.Rays(torch.from_numpy(oc).to(dev), torch.from_numpy(dc).to(dev), torch.from_numpy(dc).to(dev))
out_c = C.volume_render(ts, sv.renderer._rays_spec_from_rays(rays_c), opt)
print("  image vs explicit camera rays maxdiff:", float((out_c.view(800, 800, -1) - img).abs().max()))
print("launches:", C.launch_count())
