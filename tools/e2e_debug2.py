"""Where does the end-to-end step time go? Variants of bench.py's e2e loop (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(8, "ball"); D = 32; Q = 1 << 20
f = synth.synth_features(tr["M"], D); o, d = synth.synth_rays(Q)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
renderer = sv.VolumeRenderer(tree)
fparam = torch.from_numpy(f).to(dev).requires_grad_(True)
h_o, h_d = torch.from_numpy(o).pin_memory(), torch.from_numpy(d).pin_memory()
h_rgb, h_alpha = torch.rand(Q, 3).pin_memory(), torch.rand(Q).pin_memory()
print("pinned:", h_o.is_pinned(), h_rgb.is_pinned())
bufs = [(torch.empty(Q, 3, device=dev), torch.empty(Q, 3, device=dev), torch.empty(Q, 3, device=dev), torch.empty(Q, device=dev)) for _ in range(2)]
copy_stream = torch.cuda.Stream(device=dev); main_stream = torch.cuda.current_stream(dev)
# raw H2D bandwidth
big = torch.empty(64 << 20, dtype=torch.float32).pin_memory(); dbig = torch.empty_like(big, device=dev)
for _ in range(2): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); t = time.time()
for _ in range(5): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); print("H2D GB/s (256 MB pinned):", 5 * big.numel() * 4 / (time.time() - t) / 1e9)
free_ev = [None, None]
G3 = (D - 1) // 3
def upload(k):
    bo, bd, brgb, ba = bufs[k & 1]
    with torch.cuda.stream(copy_stream):
        if free_ev[k & 1] is not None: copy_stream.wait_event(free_ev[k & 1])
        bo.copy_(h_o, non_blocking=True); bd.copy_(h_d, non_blocking=True); brgb.copy_(h_rgb, non_blocking=True); ba.copy_(h_alpha, non_blocking=True)
        e = torch.cuda.Event(); e.record(copy_stream)
    return e
def run(n, do_upload=True, do_item=True, bump=True, simple_loss=False):
    ready = upload(0) if do_upload else None
    for k in range(n):
        nxt = upload(k + 1) if (do_upload and k + 1 < n) else None
        if do_upload: main_stream.wait_event(ready)
        bo, bd, brgb, ba = bufs[k & 1]
        fparam.grad = None
        if bump:
            with torch.no_grad(): fparam.add_(0.0)
        out = renderer(fparam, sv.Rays(bo, bd, bd))
        if simple_loss: loss = out.sum()
        else:
            rgb = out[:, :3 * G3].reshape(Q, 3, G3).mean(-1)
            loss = 0.5 * ((rgb - brgb) ** 2).mean() + 0.5 * ((out[:, -1] - ba) ** 2).mean()
        loss.backward()
        free_ev[k & 1] = torch.cuda.Event(); free_ev[k & 1].record(main_stream)
        if do_item: float(loss.item())
        ready = nxt
def timeit(label, **kw):
    run(3, **kw); torch.cuda.synchronize(); t = time.time(); run(10, **kw); torch.cuda.synchronize()
    print(f"{label:50s} {(time.time() - t) / 10 * 1e3:.2f} ms/step", flush=True)
upload(0); upload(1); torch.cuda.synchronize()
timeit("no upload, no item, no bump, sum loss", do_upload=False, do_item=False, bump=False, simple_loss=True)
timeit("no upload, no item, no bump", do_upload=False, do_item=False, bump=False)
timeit("no upload, no item, bump", do_upload=False, do_item=False)
timeit("no upload, item, bump", do_upload=False)
timeit("upload, no item, bump", do_item=False)
timeit("upload, item, bump (bench e2e)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(2, do_upload=False); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
