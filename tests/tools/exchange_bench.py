"""Leaf-gradient exchange under torchrun: the hand-written symmetric-memory kernel (NVLS multicast / peer-to-peer)
against NCCL's all-reduce -- correctness on random tables, then time per call at the C3 table size (243 MB). Dev tool.
    torchrun --nproc-per-node N tests/tools/exchange_bench.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import torch, torch.distributed as dist
from svox_t_b200 import dist as svd

rank, world, lr = svd.init_from_env("nccl")
dev = torch.device("cuda", lr)
M, D = 1897408, 32


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


ref = torch.empty(M * D, device=dev)
nccl_ms = timed(lambda: dist.all_reduce(ref))
log(f"world {world}: NCCL all-reduce of {M * D * 4 / 1e6:.0f} MB: {nccl_ms:.3f} ms, algbw {M * D * 4 / nccl_ms / 1e6:.0f} GB/s")
for backend in (os.environ.get("BACKENDS", "nvls,p2p")).split(","):
    for blocks in [int(b) for b in os.environ.get("BLOCKS", "148").split(",")]:
        try:
            x = svd.LeafGradExchange(M, D, dev, force_backend=backend, blocks=blocks)
        except Exception as e:
            log(f"{backend}: unavailable: {type(e).__name__}: {str(e)[:300]}")
            break
        log(f"{backend} blocks={blocks}: {x.describe()}")
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        worst = 0.0
        for trial in range(3):
            t = x.zeroed_table()
            t.copy_(torch.randn(M, D, device=dev, generator=g))
            want = t.clone()
            dist.all_reduce(want)
            x.all_reduce_()
            torch.cuda.synchronize()
            st = x.status()
            assert st == 0, f"exchange barrier timed out waiting for rank {st - 1}"
            err = float((x.table - want).abs().max() / want.abs().max())
            same = x.table.clone()
            dist.broadcast(same, 0)
            assert torch.equal(same, x.table), "ranks hold different sums"
            worst = max(worst, err)
        ms = timed(x.all_reduce_)
        assert x.status() == 0
        log(f"{backend} blocks={blocks}: max rel err vs NCCL {worst:.2e}; {ms:.3f} ms per call, algbw {M * D * 4 / ms / 1e6:.0f} GB/s "
            f"({nccl_ms / ms:.2f}x NCCL)")
        # rows form: 20 % of the rows are dead (sigma <= 0, identical features on every rank) and hold zeros everywhere
        feats = torch.randn(M, D, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
        feats[:, -1] = torch.rand(M, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 10 - 2
        dead = ~(feats[:, -1] > 0)
        t = x.zeroed_table()
        t.copy_(torch.randn(M, D, device=dev, generator=g))
        t[dead] = 0
        want = t.clone()
        dist.all_reduce(want)
        x.all_reduce_(features=feats)
        torch.cuda.synchronize()
        err = float((x.table - want).abs().max() / want.abs().max())
        ms_r = timed(lambda: x.all_reduce_(features=feats))
        assert x.status() == 0
        log(f"{backend} blocks={blocks} rows form ({float(dead.float().mean()) * 100:.1f} % dead rows skipped): max rel err {err:.2e}; "
            f"{ms_r:.3f} ms per call ({nccl_ms / ms_r:.2f}x NCCL)")
        del feats, want
        del x
        torch.cuda.synchronize(); dist.barrier()
dist.barrier(); dist.destroy_process_group()
