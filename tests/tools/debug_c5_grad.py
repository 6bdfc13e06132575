"""Dev tool: where does the C5-scene gradient differ from the oracle? Compares accel / reference-walk / activated paths."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
from oracle import oracle as orc
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L = int(os.environ.get("L", 10)); D = int(os.environ.get("D", 64)); Q = int(os.environ.get("Q", 4096))
tr = synth.synth_tree(L, "shell")
M = int(tr["M"])
g = torch.Generator(device=dev).manual_seed(0)
feats = torch.randn(M, D, device=dev, generator=g)
feats[:, -1] = torch.rand(M, device=dev, generator=g) * 10 - 2
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree); opt = r._get_options()
T = orc.Tree(tr["child"], tr["data"])
f_np = feats.cpu().numpy()
o, d = synth.synth_rays(Q, seed=3)
g_np = np.random.default_rng(5).standard_normal((Q, D)).astype(np.float32)
ref = torch.from_numpy(orc.render_rays_backward(T, f_np, o, d, g_np)).to(dev)
ref64 = torch.from_numpy(orc.render_rays_backward(T, f_np.astype(np.float64), o.astype(np.float64), d.astype(np.float64), g_np.astype(np.float64), dtype=np.float64).astype(np.float32)).to(dev) if os.environ.get("F64") else None
o_t, d_t, g_t = (torch.from_numpy(a).to(dev) for a in (o, d, g_np))
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
specs = {"accel_raw": tree._spec(feats), "refwalk_raw": tree._spec(feats, _with_accel=False), "accel_act_marks": r._render_spec(feats, 1 << 21)}
grads = {}
for name, spec in specs.items():
    out = C.volume_render(spec, rs, opt)
    grads[name] = C.volume_render_backward(spec, rs, opt, g_t, saved_out=out)
    G = grads[name]
    rel = float((G.double() - ref.double()).norm() / ref.double().norm())
    diff = (G - ref).abs()
    rowerr = diff.amax(dim=1)
    worst = torch.topk(rowerr, 5)
    print(f"{name}: rel L2 vs oracle {rel:.3e}; rows touched ours {int((G.abs().amax(1) > 0).sum())} oracle {int((ref.abs().amax(1) > 0).sum())}; "
          f"payload-only rel {float((G[:, :-1].double() - ref[:, :-1].double()).norm() / ref[:, :-1].double().norm()):.3e}; "
          f"sigma-only rel {float((G[:, -1].double() - ref[:, -1].double()).norm() / ref[:, -1].double().norm()):.3e}")
    for v, i in zip(worst.values.tolist(), worst.indices.tolist()):
        print(f"   row {i}: max err {v:.3e}; ours sigma-grad {float(G[i, -1]):.4e} oracle {float(ref[i, -1]):.4e}; ours |payload| {float(G[i, :-1].abs().max()):.3e} oracle {float(ref[i, :-1].abs().max()):.3e}")
    if ref64 is not None:
        print(f"   vs fp64 oracle: ours {float((G.double() - ref64.double()).norm() / ref64.double().norm()):.3e}, fp32 oracle {float((ref.double() - ref64.double()).norm() / ref64.double().norm()):.3e}")
names = list(grads)
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        a, b = grads[names[i]], grads[names[j]]
        print(f"{names[i]} vs {names[j]}: rel {float((a.double() - b.double()).norm() / b.double().norm()):.3e}")
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import refdrv
if refdrv.available():
    m = refdrv.module()
    rts = refdrv.tree_spec(feats, tree.child, tree.data, tree.parent_depth, tree.offset, tree.invradius, tree.filled)
    rrs, ro = refdrv.rays_spec(o_t, d_t), refdrv.options()
    rg = m.volume_render_backward(rts, rrs, ro, g_t)
    ro_out = m.volume_render(rts, rrs, ro)
    mine = grads["accel_act_marks"]
    print(f"LIVE REFERENCE: ours vs ref rel {float((mine.double() - rg.double()).norm() / rg.double().norm()):.3e}; "
          f"oracle vs ref rel {float((ref.double() - rg.double()).norm() / rg.double().norm()):.3e}; rows touched ref {int((rg.abs().amax(1) > 0).sum())}")
    out = C.volume_render(specs["accel_act_marks"], rs, opt)
    oo = torch.from_numpy(orc.render_rays(T, f_np, o, d)[0]).to(dev)
    print(f"fwd: ours vs ref max {float((out - ro_out).abs().max()):.3e}; oracle vs ref max {float((oo - ro_out).abs().max()):.3e}; "
          f"rays differing > 1e-5: ours {int(((out - ro_out).abs().amax(1) > 1e-5).sum())} oracle {int(((oo - ro_out).abs().amax(1) > 1e-5).sum())}")
