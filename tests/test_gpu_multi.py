"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): rays sharded over the ranks, tree + features
replicated, NCCL all-reduce of the leaf gradients == the single-GPU gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import svox_t_b200 as sv
    from svox_t_b200 import dist as svd, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    svd.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    tr = synth.synth_tree(6, "ball")
    D, Q = 32, 40000
    f = synth.synth_features(tr["M"], D)
    o, d = synth.synth_rays(Q)
    g = np.random.default_rng(5).standard_normal((Q, D)).astype(np.float32)
    tree = sv.N3Tree.from_tensors(tr["child"], tr["data"], tr["parent_depth"], data_dim=D, map_location=dev)
    feats = torch.from_numpy(f).to(dev).requires_grad_(True)
    rays = sv.Rays(*(torch.from_numpy(a).to(dev) for a in (o, d, d)))
    g_t = torch.from_numpy(g).to(dev)
    r = sv.VolumeRenderer(tree)
    loss, grad = svd.render_step_sharded(r, feats, rays, lambda out, lo, hi: (out * g_t[lo:hi]).sum(), rank, world)
    torch.cuda.synchronize()
    # one frame sharded by row bands, gathered on every rank == the frame rendered by one GPU
    cam = torch.from_numpy(synth.synth_cameras(1)[0]).to(dev)
    frame = svd.render_image_bands(r, feats.detach(), cam, 90, 61, 100.0, rank=rank, world=world)
    frame_ok = bool(torch.equal(frame, r.render_persp(feats.detach(), cam, width=90, height=61, fx=100.0)))
    imgs, (lo, hi) = svd.render_views_sharded(r, feats.detach(), [cam] * 3, 32, 24, 40.0, rank=rank, world=world)
    assert len(imgs) == hi - lo and frame_ok
    # the same step with the hand-written exchange (symmetric memory; NVLS multicast or peer-to-peer): the backward
    # reduces into the exchange's table, the sum over the GPUs happens inside backward()
    xchg = svd.LeafGradExchange(tr["M"], D, dev)
    r.leaf_grad_exchange = xchg
    f2 = feats.detach().clone().requires_grad_(True)
    lo, hi = svd.shard_range(Q, rank, world)
    (r(f2, svd.shard_rays(rays, rank, world)) * g_t[lo:hi]).sum().backward()
    torch.cuda.synchronize()
    assert xchg.status() == 0
    # gradient accumulation over two backward calls: the first sum must survive the reuse of the exchange's table
    f3 = feats.detach().clone().requires_grad_(True)
    mine = svd.shard_rays(rays, rank, world)
    (r(f3, mine) * g_t[lo:hi]).sum().backward()
    (r(f3, mine) * g_t[lo:hi]).sum().backward()
    torch.cuda.synchronize()
    acc_rel = float((f3.grad - 2 * f2.grad).norm() / f2.grad.norm())
    r.leaf_grad_exchange = None
    if rank == 0:
        full = feats.detach().clone().requires_grad_(True)
        (r(full, rays) * g_t).sum().backward()
        rel = float((grad - full.grad).norm() / full.grad.norm())
        rel2 = float((f2.grad - full.grad).norm() / full.grad.norm())
        q.put((rel, rel2, xchg.describe()["backend"], acc_rel))
    svd.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_gradients_match_single_gpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rel, rel2, backend, acc_rel = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert rel < 1e-5, rel
    assert rel2 < 1e-5, (rel2, backend)
    assert acc_rel < 1e-5, acc_rel
