"""Host-side ceiling of the e2e leg: plain cudaMemcpyAsync bandwidth from page-locked host memory (what the copy engine reaches
with large requests) beside what SM-issued reads of scattered 32-byte sectors reach (tools/e2e_step.py).  One process per GPU under
torchrun measures the concurrent case (all ranks copy at once): python -m torch.distributed.run --nproc-per-node N tools/h2d_ceiling.py"""
import os
import time

import torch

rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("rank %d of %d: pinned H2D cudaMemcpyAsync 1 GiB: %.1f GB/s" % (rank, world, h.numel() / dt / 1e9), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
