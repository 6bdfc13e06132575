"""How fast are the march kernels when the feature / gradient tables fit in L2?  Same leaf size (depth 8), same ray
recipe, smaller balls: r = 0.30 (C3: 243 MB table), 0.20, 0.15 (30 MB), 0.10. Prints ms and ns per hit sample."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
from oracle import oracle as orc

Q = 1 << 20
D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
for r in (0.30, 0.20, 0.15, 0.10):
    tr = synth.synth_tree(8, "ball", r_out=r)
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q, r_target=r)
    T = orc.Tree(tr["child"], tr["data"])
    cnt = orc.render_rays(T, f, o[:4096], d[:4096], want_counters=True)[2]
    S, V, H = (cnt[k] / cnt["Q"] for k in ("S", "V", "H"))
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    feats = torch.from_numpy(f).to(dev)
    o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
    rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t))
    opt = sv.VolumeRenderer(tree)._get_options()
    ts = tree._spec(feats)
    g = torch.randn(Q, D, device=dev)
    best = [1e9, 1e9]
    for _ in range(5):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); out = C.volume_render(ts, rs, opt); e1.record()
        grad = C.volume_render_backward(ts, rs, opt, g, saved_out=out); e2.record()
        torch.cuda.synchronize()
        best = [min(best[0], e0.elapsed_time(e1)), min(best[1], e1.elapsed_time(e2))]
    print(f"r={r:.2f} M={tr['M']} table {tr['M']*D*4/1e6:.0f} MB  S={S:.1f} V={V:.1f} H={H:.1f}  fwd {best[0]:.3f} ms "
          f"({best[0]*1e6/(Q*V):.3f} ns/valid sample)  bwd {best[1]:.3f} ms ({best[1]*1e6/(Q*V):.3f} ns/valid sample)", flush=True)
