"""Aggregate device -> host copy ceiling of the box: every rank copies the e2e leg's per-step result volume (316 MB, pinned
destination) at the same time, first rank 0 alone, then all ranks together.  Explains `e2e` at N = 8.
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/d2h_ceiling.py"""
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
NBYTES = 316 << 20
src = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
dst = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()


def run(active, reps=10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    if active:
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item(), reps * NBYTES / ms / 1e6 if active else 0.0


run(True, 2)
ms1, gbs1 = run(rank == 0)
msn, gbsn = run(True)
g = torch.tensor([gbsn], device="cuda")
if world > 1:
    dist.all_reduce(g)
if rank == 0:
    print(f"rank 0 alone: {gbs1:.1f} GB/s;  {world} ranks together: {g.item():.1f} GB/s aggregate, slowest rank took "
          f"{msn / 10:.2f} ms per 316 MB")
if world > 1:
    dist.destroy_process_group()
