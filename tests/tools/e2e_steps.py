"""Per-step wall time vs GPU time of the e2e loop (dev tool)."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(8, "ball"); D = 32; Q = 1 << 20
f = synth.synth_features(tr["M"], D); o, d = synth.synth_rays(Q)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
renderer = sv.VolumeRenderer(tree)
fparam = torch.from_numpy(f).to(dev).requires_grad_(True)
bo, bd = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
brgb, ba = torch.rand(Q, 3, device=dev), torch.rand(Q, device=dev)
G3 = (D - 1) // 3
mode = sys.argv[1] if len(sys.argv) > 1 else "full"
rows = []
for k in range(70):
    t0 = time.time()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    fparam.grad = None
    e[0].record()
    if mode != "nobump":
        with torch.no_grad(): fparam.add_(0.0)
    out = renderer(fparam, sv.Rays(bo, bd, bd))
    e[1].record()
    rgb = out[:, :3 * G3].reshape(Q, 3, G3).mean(-1)
    loss = 0.5 * ((rgb - brgb) ** 2).mean() + 0.5 * ((out[:, -1] - ba) ** 2).mean()
    e[2].record()
    loss.backward()
    e[3].record()
    t1 = time.time()
    v = float(loss.item())
    t2 = time.time()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t0), e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3])))
    if k % 10 == 9:
        smi = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        print("step", k, "smi:", smi, flush=True)
for k, r in enumerate(rows):
    print(f"{k:3d} cpu-launch {r[0]:6.2f} wall {r[1]:6.2f} | gpu fwd {r[2]:6.2f} loss {r[3]:5.2f} bwd {r[4]:6.2f}")
