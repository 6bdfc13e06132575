"""Config C5 on N GPUs (BASELINE.json configs[4]): depth-10 shell octree (~33 M rows x 64 ch = 8.4 GB features), tree and
features replicated on every GPU, a batch of V 1920x1080 views split across the ranks by whole views
(dist.render_views_sharded). Launch: torchrun --nproc-per-node N tests/tools/scene_c5_dist.py [V].
Rank 0 builds the synthetic tree on the host and broadcasts its tensors; features are generated per rank from one seed.
Timing: CUDA events per rank around the rank's share, max over ranks (all_reduce MAX), after one warm-up pass."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import torch.distributed as td
import svox_t_b200 as sv
from svox_t_b200 import synth, dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    td.init_process_group("nccl", device_id=dev)
V = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * world
L, D, W, H, FX = 10, 64, 1920, 1080, 1500.0

t0 = time.time()
if rank == 0:
    tr = synth.synth_tree(L, "shell")
    shapes = torch.tensor([tr["n_nodes"], tr["M"]], dtype=torch.int64, device=dev)
else:
    shapes = torch.zeros(2, dtype=torch.int64, device=dev)
if world > 1:
    td.broadcast(shapes, 0)
n, M = int(shapes[0]), int(shapes[1])
if rank == 0:
    child, data, pd = (torch.from_numpy(tr[k]).to(dev) for k in ("child", "data", "parent_depth"))
else:
    child = torch.empty((n, 2, 2, 2), dtype=torch.int32, device=dev)
    data = torch.empty((n, 2, 2, 2, 1), dtype=torch.int32, device=dev)
    pd = torch.empty((n, 2), dtype=torch.int32, device=dev)
if world > 1:
    for t in (child, data, pd):
        td.broadcast(t, 0)
g = torch.Generator(device=dev); g.manual_seed(0)
feats = torch.randn(M, D, device=dev, generator=g)
feats[:, -1] = torch.rand(M, device=dev, generator=g) * 10 - 2
tree = sv.N3Tree.from_tensors(child, data, pd, data_dim=D, map_location=dev)
r = sv.VolumeRenderer(tree)
cams = [torch.from_numpy(c).to(dev) for c in synth.synth_cameras(V)]
setup_s = time.time() - t0

def one_pass():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()
    a.record()
    imgs, (lo, hi) = dist.render_views_sharded(r, feats, cams, W, H, FX, rank=rank, world=world)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device=dev)
    hit = torch.tensor([sum(float((im[..., -1] > 0).float().mean()) for im in imgs), float(hi - lo)], device=dev)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
        td.all_reduce(hit)
    return float(ms), float(hit[0] / hit[1])

one_pass()
res = [one_pass() for _ in range(3)]
ms = float(np.median([x[0] for x in res]))
if rank == 0:
    print(json.dumps({"config": "C5", "n_gpus": world, "views": V, "width": W, "height": H, "rows": M, "D": D, "nodes": n,
                      "ms_per_batch_max_over_ranks": round(ms, 3), "ms_per_view_per_gpu": round(ms / (V / world), 3),
                      "Mpixel_per_s_all_gpus": round(V * W * H / ms / 1e3, 1), "hit_fraction": round(res[0][1], 3),
                      "features_GB_per_gpu": round(M * D * 4 / 1e9, 2), "setup_s": round(setup_s, 1),
                      "accel": tree.accel(feats).describe()}), flush=True)
if world > 1:
    td.destroy_process_group()
