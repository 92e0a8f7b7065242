"""Multi-GPU plumbing: video streams are independent, so the path shards by
stream with NO data-path collective.  One process per GPU (torchrun); torch.distributed is used only
for the barrier around the timed region and the max-over-ranks of the per-rank timings."""
from __future__ import annotations

from typing import List, Sequence


def shard_streams(n_streams: int, world: int, rank: int) -> List[int]:
    """Contiguous block partition of stream ids over ranks (32 streams: 4/GPU at 8 GPUs, 8/GPU at 4, ...).
    Remainders go to the lowest ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_streams, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def max_over_ranks(values: Sequence[float], device="cpu") -> List[float]:
    """Element-wise MAX of per-rank timings over the default process group (NCCL on GPUs, gloo in
    the CPU tests); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def aggregate_throughput(units_per_rank: int, world: int, steps: int, max_total_ms: float) -> float:
    """Whole-job units/s: all ranks' units over the slowest rank's time."""
    return units_per_rank * world * steps / (max_total_ms / 1e3)
