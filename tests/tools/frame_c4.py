"""Config C4 (BASELINE.json): animated frame = LBS warp of 2^20 voxel centres + p2v splat (256^3) + octree rebuild
to depth 8 from the warped points + 1920x1080 render with opacity and depth. Per-stage and total latency."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C

dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L, D, P = 8, 32, 1 << 20
rng = np.random.default_rng(2)
vox = synth._occupied_keys(L, "ball")
pts = synth.voxel_centers(vox[rng.permutation(len(vox))[:P]], L)
Tm, w, ji = synth.synth_skeleton(P)
f = synth.synth_features(P, D)
p, Tm_t, w_t, ji_t, feats = (torch.from_numpy(a).to(dev) for a in (pts, Tm, w, ji, f))
corner, size = torch.zeros(3, device=dev), torch.ones(3, device=dev)
cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
tree = sv.N3Tree(N=2, data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree)

def frame(rebuild="oneshot"):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    warped, mats = sv.warp_vertices(Tm_t, p, w_t, ji_t)
    ev[1].record()
    grid = sv.voxelize(warped, feats, corner, size, 256, 1.5 / 256, 2.0 / 256)
    ev[2].record()
    if rebuild == "oneshot":
        tree.build_from_points(warped, L)
    elif rebuild == "oneshot_nosync":
        tree.build_from_points(warped, L, capacity=CAP)
    else:
        t2 = sv.N3Tree(N=2, data_dim=D, init_reserve=300000, map_location=dev)
        for _ in range(L - 1):
            t2[warped].refine()
        t2.construct_tree(warped)
        tree.child, tree.data, tree.parent_depth, tree.filled = t2.child, t2.data, t2.parent_depth, t2.filled
        tree._invalidate()
    ev[3].record()
    acc = tree.accel(feats)
    ev[4].record()
    img, depth = r.render_persp_with_depth(feats, cam, width=1920, height=1080, fx=1500.0)
    ev[5].record()
    torch.cuda.synchronize()
    names = ["warp_vertices", "p2v", "rebuild", "accel", "render_1080p"]
    t = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
    t["total"] = ev[0].elapsed_time(ev[5])
    return t, img, depth, grid

tree.build_from_points(sv.warp_vertices(Tm_t, p, w_t, ji_t)[0], L)
CAP = int(tree.filled * 1.25)
for mode in ("oneshot_nosync", "oneshot", "refine_loop"):
    for _ in range(3):
        t, img, depth, grid = frame(mode)
    ts = [frame(mode)[0] for _ in range(5)]
    med = {k: float(np.median([x[k] for x in ts])) for k in ts[0]}
    print(mode, json.dumps({k: round(v, 3) for k, v in med.items()}), "nodes", tree.filled,
          "hit frac", float((img[..., -1] > 0).float().mean()), "grid sum", float(grid.sum()))
try:
    import refdrv
    if refdrv.available():
        m = refdrv.module()
        def best(fn, n=5):
            out = []
            for _ in range(n):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize(); out.append(a.elapsed_time(b))
            return float(np.median(out))
        print("REF warp_vertices ms", best(lambda: m.warp_vertices(Tm_t, p, w_t, ji_t)))
        warped = m.warp_vertices(Tm_t, p, w_t, ji_t)[0]
        print("REF p2v ms", best(lambda: m.p2v(warped, feats, corner, size, 256, 1.5 / 256, 2.0 / 256)))
        mine = sv.warp_vertices(Tm_t, p, w_t, ji_t)[0]
        print("warp maxdiff vs REF", float((mine - warped).abs().max()),
              "p2v maxdiff vs REF", float((sv.voxelize(warped, feats, corner, size, 256, 1.5/256, 2.0/256) - m.p2v(warped, feats, corner, size, 256, 1.5/256, 2.0/256)).abs().max()))
except Exception as e:
    print("ref compare skipped:", e)
