"""Host<->device copy bandwidth per rank, alone and with all ranks copying at once (torchrun, one rank per GPU).
Answers whether the end-to-end path at N GPUs is bounded by the box's shared PCIe / host-memory bandwidth:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py
Prints GB/s per rank for H2D, D2H and both directions at once, (a) rank by rank, (b) all ranks together."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pin = "--pin" in sys.argv
info = bench.pin_rank(rank, world, local, "on" if pin else "off")
dev = torch.device("cuda", local)
GB = 1.2
n = int(GB * 2**30 / 4)
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_in.fill_(1.0)
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
h_out.fill_(0.0)
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.ones(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(mode, reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    per_dir = GB * 2**30 * reps / dt / 1e9
    return per_dir


def barrier():
    if world > 1:
        dist.barrier()


for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    # alone: rank by rank
    alone = None
    for r in range(world):
        barrier()
        if r == rank:
            alone = run(mode)
    barrier()
    together = run(mode)
    barrier()
    res = [None] * world
    if world > 1:
        dist.all_gather_object(res, (rank, round(alone, 1), round(together, 1), info["cpus"]))
    else:
        res = [(rank, round(alone, 1), round(together, 1), info["cpus"])]
    if rank == 0:
        print(f"{mode:5s} GB/s per direction  alone: {[r[1] for r in res]}  all ranks at once: {[r[2] for r in res]}  "
              f"sum {sum(r[2] for r in res):.0f}  (pin={pin}, cpus/rank {res[0][3]})", flush=True)
if world > 1:
    dist.destroy_process_group()
