"""How long does ONE frame's pinned host->device copy (8.2 MB = 2000 fp16 rows x 2048 + boxes + scores) take when
N ranks of one node copy at the same time?  The floor of bench.py's end-to-end step at N GPUs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_scaling_probe.py
Rank 0 prints one line: per-copy time (max over ranks of each rank's mean over 200 back-to-back barrier-aligned
copies) and the aggregate bandwidth."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 8_232_000
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(5):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
iters = 200
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
ms = 1e3 * (time.perf_counter() - t0) / iters
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"{world} rank(s): {float(t):.4f} ms per 8.2 MB pinned H2D copy (slowest rank), "
          f"{world * nbytes / (float(t) * 1e-3) / 1e9:.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
