"""A few C4 frames in one rebuild mode, for the ncu launch list (kernel shares of the frame). Dev tool.
    python tests/tools/frame_c4_once.py [bitmap|sort] [frames]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
mode = sys.argv[1] if len(sys.argv) > 1 else "bitmap"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
L, D, P = 8, 32, 1 << 20
vox = synth._occupied_keys(L, "ball")
pts = synth.voxel_centers(vox[np.random.default_rng(2).permutation(len(vox))[:P]], L)
Tm, w, ji = synth.synth_skeleton(P)
p, Tm_t, w_t, ji_t, feats = (torch.from_numpy(a).to(dev) for a in (pts, Tm, w, ji, synth.synth_features(P, D)))
corner, size = torch.zeros(3, device=dev), torch.ones(3, device=dev)
cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
tree = sv.N3Tree(N=2, data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree)
tree.build_from_points(sv.warp_vertices(Tm_t, p, w_t, ji_t)[0], L)
cap = int(tree.filled * 1.25)
torch.cuda.synchronize()
for f in range(frames):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    warped, _ = sv.warp_vertices(Tm_t, p, w_t, ji_t)
    ev[1].record()
    sv.voxelize(warped, feats, corner, size, 256, 1.5 / 256, 2.0 / 256)
    ev[2].record()
    if mode == "bitmap":
        tree.build_from_points(warped, L, capacity=cap)
    else:
        c, d_, pd, _ = sv.csrc.build_octree(warped, L, tree.offset, tree.invradius, sort_based=True)
        tree.child, tree.data, tree.parent_depth, tree.filled = c, d_, pd, c.shape[0]
        tree._invalidate(); tree._known_depth = L
    ev[3].record()
    tree.accel(feats)
    ev[4].record()
    r.render_persp_with_depth(feats, cam, width=1920, height=1080, fx=1500.0)
    ev[5].record()
    torch.cuda.synchronize()
    print(mode, "frame", f, [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(5)], "total", round(ev[0].elapsed_time(ev[5]), 3), flush=True)
