"""probe (torchrun, one rank per GPU): raw pinned-host -> device copy rate of 96 MB per rank with all ranks copying at once -- the floor of
the end-to-end numbers at N GPUs (what the box's host memory / PCIe fabric delivers, no kernels involved)"""
import os, sys, time
import torch, torch.distributed as dist
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 12_000_000
h = torch.empty(n, dtype=torch.float64).pin_memory(); h.uniform_(-3, 3)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for it in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for it in range(40): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t = (time.perf_counter() - t0) / 40
x = torch.tensor([t], dtype=torch.float64, device="cuda")
if world > 1: dist.all_reduce(x, op=dist.ReduceOp.MAX)
if rank == 0:
    print("H2D 96 MB per rank, %d ranks at once: slowest rank %.3f ms = %.1f GB/s per rank, %.1f GB/s aggregate = %.3e S-GD samples/s at 96 B each" % (
        world, 1e3 * x.item(), 0.096 / x.item(), world * 0.096 / x.item(), world * 1e6 / x.item()))
if world > 1: dist.destroy_process_group()
