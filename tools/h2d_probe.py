"""torchrun --nproc-per-node G tools/h2d_probe.py : pinned host -> device bandwidth of every rank at once.
The e2e number of bench.py moves 9.4 GB host -> device per step whatever G is (the bit plane of V is cut into column
blocks, one per rank); this measures what the box gives G concurrent uploads, i.e. the floor of that number."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else int(1.2e9)
h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); h.fill_(1)
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
d2 = torch.empty(nbytes // 4, dtype=torch.uint8, device="cuda")
h2 = torch.empty(nbytes // 4, dtype=torch.uint8, pin_memory=True)
for what in ("h2d", "h2d+d2h"):
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s2 = torch.cuda.Stream()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
        if what != "h2d":
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = 5 * nbytes / t.item() / 1e9
        print(f"{world} rank(s), {what}: {gbs:.1f} GB/s host->device per rank, {gbs * world:.1f} GB/s aggregate "
              f"({nbytes / 1e9:.2f} GB per copy, pinned, slowest rank)", flush=True)
if world > 1:
    dist.destroy_process_group()
