import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import svox_t_b200 as sv
from svox_t_b200 import synth, csrc as C
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
tr = synth.synth_tree(8, "ball"); D=32; Q=1<<20
f = synth.synth_features(tr["M"], D); o, d = synth.synth_rays(Q)
tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
feats = torch.from_numpy(f).to(dev)
renderer = sv.VolumeRenderer(tree)
fparam = feats.clone().requires_grad_(True)
o_t, d_t = torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev)
tgt = torch.rand(Q, D, device=dev)
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t=time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time()-t)/n*1e3
def full(chunks):
    cq = Q // chunks
    fparam.grad = None
    tot = torch.zeros((), device=dev)
    for c in range(chunks):
        sl = slice(c*cq, (c+1)*cq)
        out = renderer(fparam, sv.Rays(o_t[sl], d_t[sl], d_t[sl]))
        loss = 0.5 * ((out - tgt[sl]) ** 2).sum() / (Q * D)
        loss.backward(); tot += loss.detach()
    return float(tot.item())
for ch in (1, 2, 4):
    print("device-resident autograd step, chunks", ch, "ms", ev(lambda: full(ch)))
def fwd_only():
    with torch.no_grad(): return renderer(fparam, sv.Rays(o_t, d_t, d_t))
print("fwd via API ms", ev(fwd_only))
out = fwd_only()
def loss_only():
    o2 = out.clone().requires_grad_(True)
    l = 0.5 * ((o2 - tgt) ** 2).sum() / (Q * D); l.backward(); return l
print("loss fwd+bwd elementwise ms", ev(loss_only))
ts = renderer._render_spec(fparam.detach(), Q)
rs = sv.renderer._rays_spec_from_rays(sv.Rays(o_t, d_t, d_t)); opt = renderer._get_options()
g1 = torch.randn(Q, D, device=dev)
g2 = ((out - tgt) / (Q * D)).contiguous()
print("bwd randn g ms", ev(lambda: C.volume_render_backward(ts, rs, opt, g1, saved_out=out)))
print("bwd loss g ms", ev(lambda: C.volume_render_backward(ts, rs, opt, g2, saved_out=out)))
print("bwd loss g*1e6 ms", ev(lambda: C.volume_render_backward(ts, rs, opt, (g2*1e6).contiguous(), saved_out=out)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    full(1); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
