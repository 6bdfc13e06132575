"""All-reduce of the leaf-gradient table (243 MB fp32) under torchrun: time per call, algorithm bandwidth (dev tool).
    torchrun --nproc-per-node N tests/tools/allreduce_bench.py            (NCCL_ALGO / NCCL_DEBUG from the environment)"""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 1897408 * 32
x = torch.ones(n, device="cuda")
for _ in range(5): dist.all_reduce(x)
torch.cuda.synchronize(); dist.barrier()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dist.all_reduce(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
if rank == 0:
    ms = ts[len(ts) // 2]
    print(f"world {world} NCCL_ALGO={os.environ.get('NCCL_ALGO', 'default')}: all-reduce of {n * 4 / 1e6:.0f} MB median {ms:.3f} ms, "
          f"algbw {n * 4 / ms / 1e6:.0f} GB/s, busbw {n * 4 / ms / 1e6 * 2 * (world - 1) / world:.0f} GB/s", flush=True)
dist.barrier(); dist.destroy_process_group()
