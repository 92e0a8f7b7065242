"""Multi-GPU plumbing: video streams are independent, so the path shards by
stream and no collective exists INSIDE the algorithm.  One process per GPU (torchrun, NCCL); the only
communication is the per-stream scatter of a frame's detector outputs from the producer rank to the ranks that
own the streams and the gather of their results (BASELINE config 5: "NCCL only for the trivial per-stream
scatter"), plus the barrier / max-over-ranks of the timings.  Everything here works on whatever backend the
default process group uses (NCCL + device tensors on the GPUs, gloo + CPU tensors in the tests)."""
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


# ---- host placement of the ranks (one process per GPU on one node) -------------------------------------------
def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist) -> [0, 1, 2, 3, 8, 10, 11]."""
    out: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.extend(range(int(lo), int(hi or lo) + 1))
    return out


def split_cores_numa_local(local_cpus: Sequence[Sequence[int]], allowed: Sequence[int], local_rank: int) -> List[int]:
    """Host cores for rank `local_rank`: `local_cpus[r]` = the cores next to rank r's GPU (its PCIe root's NUMA
    node; empty when unknown).  Ranks whose GPUs hang off the same node share that node's cores in equal
    contiguous slices; a rank with no topology information falls back to an equal slice of `allowed`.  The
    rank's pinned staging buffers are then first-touched on the node its GPU's DMA reads from: without this,
    8 ranks' 8 MB frame copies cross the socket interconnect and the end-to-end step time grows with the
    number of ranks although no rank talks to another."""
    allowed = sorted(allowed)
    world = len(local_cpus)
    mine = sorted(set(local_cpus[local_rank]) & set(allowed))
    if not mine:
        per = max(1, len(allowed) // max(1, world))
        return allowed[local_rank * per:(local_rank + 1) * per] or allowed
    peers = [r for r in range(world) if sorted(set(local_cpus[r]) & set(allowed)) == mine]
    per = max(1, len(mine) // len(peers))
    i = peers.index(local_rank)
    return mine[i * per:(i + 1) * per] or mine


def gpu_local_cpus(device_index: int) -> List[int]:
    """Cores of the NUMA node GPU `device_index` is attached to (sysfs `local_cpulist` of its PCI function),
    [] when the topology is not exposed."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            return parse_cpulist(f.read())
    except Exception:  # noqa: BLE001 -- no sysfs / old torch: the caller falls back to an even split
        return []


# ---- per-stream scatter / gather (BASELINE config 5) -------------------------------------------------------
# One frame of one video stream travels as a fixed-size int32 slot (so that a rank's streams are one contiguous
# message and `dist.scatter` needs no size negotiation):
#     [m | boxes int32[4*cap] | scores float32 bits [cap] | gt int32[cap]]
# and its result as
#     [n | ids int32[cap] | tlbr float64 bits [8*cap]]
def frame_slot_ints(cap: int) -> int:
    return 1 + 6 * cap


def result_slot_ints(cap: int) -> int:
    return 1 + 9 * cap


def pack_frame(boxes, scores, gt, cap: int):
    """boxes int32[m,4], scores float32[m], gt int[m] -> int32[frame_slot_ints(cap)] (NumPy)."""
    import numpy as np
    m = int(len(boxes))
    if m > cap:
        raise ValueError(f"{m} detections exceed the slot capacity {cap}")
    out = np.zeros(frame_slot_ints(cap), np.int32)
    out[0] = m
    out[1:1 + 4 * m] = np.asarray(boxes, np.int32).reshape(-1)
    out[1 + 4 * cap:1 + 4 * cap + m] = np.asarray(scores, np.float32).view(np.int32)
    out[1 + 5 * cap:1 + 5 * cap + m] = np.asarray(gt).astype(np.int32)
    return out


def unpack_frame(slot, cap: int):
    """Inverse of pack_frame on a torch or NumPy int32 vector: (m, boxes[m,4], scores[m] float32, gt[m])."""
    m = int(slot[0])
    boxes = slot[1:1 + 4 * m].reshape(m, 4)
    sc_bits = slot[1 + 4 * cap:1 + 4 * cap + m]
    scores = None
    try:
        import torch
        if isinstance(slot, torch.Tensor):
            scores = sc_bits.view(torch.float32)
    except ImportError:      # pragma: no cover
        pass
    if scores is None:
        import numpy as np
        scores = np.asarray(sc_bits).view(np.float32)
    gt = slot[1 + 5 * cap:1 + 5 * cap + m]
    return m, boxes, scores, gt


def pack_result(ids, tlbr, cap: int):
    import numpy as np
    n = int(len(ids))
    out = np.zeros(result_slot_ints(cap), np.int32)
    out[0] = n
    out[1:1 + n] = np.asarray(ids, np.int32)
    out[1 + cap:1 + cap + 8 * n] = np.ascontiguousarray(tlbr, np.float64).reshape(-1).view(np.int32)
    return out


def unpack_result(slot, cap: int):
    import numpy as np
    slot = np.asarray(slot)
    n = int(slot[0])
    ids = slot[1:1 + n].copy()
    tlbr = slot[1 + cap:1 + cap + 8 * n].copy().view(np.float64).reshape(n, 4)
    return ids, tlbr


def scatter_streams(packed_all, n_streams: int, slot_ints: int, src: int = 0, device="cpu"):
    """Per-stream scatter: `packed_all` (rank `src` only; int32 tensor [n_streams, slot_ints] on `device`) is cut
    into the contiguous stream blocks of shard_streams() and scattered; returns this rank's [my_streams, slot_ints]."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    mine = shard_streams(n_streams, world, rank)
    if world == 1:
        return packed_all
    recv = torch.empty((len(mine), slot_ints), dtype=torch.int32, device=device)
    if rank == src:
        parts = []
        for r in range(world):
            ids = shard_streams(n_streams, world, r)
            parts.append(packed_all[ids[0]:ids[-1] + 1].contiguous() if ids else torch.empty((0, slot_ints), dtype=torch.int32, device=device))
        sizes = {p.shape[0] for p in parts}
        if len(sizes) == 1:
            dist.scatter(recv, parts, src=src)
        else:                     # ragged partition: point-to-point
            reqs = [dist.isend(parts[r], dst=r) for r in range(world) if r != src and parts[r].numel()]
            recv.copy_(parts[src])
            for q in reqs:
                q.wait()
    else:
        sizes_equal = n_streams % world == 0
        if sizes_equal:
            dist.scatter(recv, None, src=src)
        elif recv.numel():
            dist.recv(recv, src=src)
    return recv


def gather_streams(mine_packed, n_streams: int, slot_ints: int, dst: int = 0, device="cpu"):
    """Inverse: every rank's [my_streams, slot_ints] results are gathered to rank `dst` as [n_streams, slot_ints]
    (None elsewhere)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return mine_packed
    if n_streams % world == 0:
        outs = [torch.empty_like(mine_packed) for _ in range(world)] if rank == dst else None
        dist.gather(mine_packed, outs, dst=dst)
        return torch.cat(outs, 0) if rank == dst else None
    if rank == dst:
        outs = []
        for r in range(world):
            ids = shard_streams(n_streams, world, r)
            buf = torch.empty((len(ids), slot_ints), dtype=torch.int32, device=device)
            if r == dst:
                buf.copy_(mine_packed)
            elif buf.numel():
                dist.recv(buf, src=r)
            outs.append(buf)
        return torch.cat(outs, 0)
    if mine_packed.numel():
        dist.send(mine_packed, dst=dst)
    return None
