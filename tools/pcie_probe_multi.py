"""Aggregate host<->device copy ceiling with every rank copying at once (pinned memory, both directions).
   torchrun --nproc-per-node N tools/pcie_probe_multi.py   -> one line per mode, summed over ranks.
Explains the end-to-end leg of bench.py at N > 1: on this pool's host the aggregate stops growing after two GPUs."""
import os, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 398131200                                   # one 64-frame 1080p batch
a = torch.empty(n, dtype=torch.uint8).pin_memory(); b = torch.empty(n, dtype=torch.uint8).pin_memory()
a.fill_(1); b.fill_(2)
da = torch.empty(n, dtype=torch.uint8, device="cuda"); db = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item()) / reps


run(1, 1, 2)
for name, h, d in (("h2d only", 1, 0), ("d2h only", 0, 1), ("both directions", 1, 1)):
    t = run(h, d)
    if rank == 0:
        print(f"N={world} {name}: {n / t / 1e9:.1f} GB/s per direction per GPU, {world * n / t / 1e9:.1f} GB/s per direction aggregate"
              f" -> {world * 64 / t:.0f} frames/s ceiling" if (h and d) else
              f"N={world} {name}: {n / t / 1e9:.1f} GB/s per GPU, {world * n / t / 1e9:.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
