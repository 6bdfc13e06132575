"""World-size-2 gloo test of the multi-GPU host logic on CPU: rays shard without overlap, and the all-reduced
per-rank leaf gradients equal the single-process gradient. The per-rank 'kernel' here is the CPU oracle (tests may
use it); on the B200 box the same dist.py functions drive the CUDA path over NCCL (tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from svox_t_b200 import dist as svd
from svox_t_b200 import synth


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            spans = [svd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [svd.shard_image_rows(1080, r, 8) for r in range(8)][-1][1] == 1080


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from oracle import oracle as orc
    from svox_t_b200.renderer import Rays
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    svd.init_from_env("gloo")
    tr = synth.synth_tree(3, "ball", r_out=0.45)
    T = orc.Tree(tr["child"], tr["data"])
    f = synth.synth_features(tr["M"], 6)
    o, d = synth.synth_rays(301)
    g = np.random.default_rng(5).standard_normal((301, 6)).astype(np.float32)
    rays = Rays(torch.from_numpy(o), torch.from_numpy(d), torch.from_numpy(d))
    mine = svd.shard_rays(rays, rank, world)
    lo, hi = svd.shard_range(301, rank, world)
    grad = torch.from_numpy(orc.render_rays_backward(T, f, mine.origins.numpy(), mine.dirs.numpy(), g[lo:hi]))
    svd.all_reduce_leaf_grads(grad)
    t_ms = svd.max_over_ranks(10.0 * (rank + 1), torch.device("cpu"))
    svd.barrier()
    if rank == 0:
        full = orc.render_rays_backward(T, f, o, d, g)
        q.put((float(np.linalg.norm(grad.numpy() - full) / np.linalg.norm(full)), t_ms, hi - lo))
    dist.destroy_process_group()


def test_two_rank_gradient_sum_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rel, t_ms, n0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rel < 1e-5 and t_ms == 20.0 and n0 == 151
