"""N>1 path on CPU: world_size-2 gloo run of the stream sharding and the max-over-ranks reduction
that bench.py uses (no data-path collective exists on this path)."""
import os
import socket

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from botsort_b200.sharding import aggregate_throughput, max_over_ranks, shard_streams


def test_shard_streams_partitions_everything():
    for n, world in ((32, 8), (32, 4), (32, 2), (32, 1), (5, 2), (3, 4)):
        parts = [shard_streams(n, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard_streams(32, 8, 3) == [12, 13, 14, 15]
    with pytest.raises(ValueError):
        shard_streams(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_streams(6, world, rank)
    local_ms = [10.0 + 5.0 * rank, 1.0 + rank]          # rank 1 is the slow one
    red = max_over_ranks(local_ms)
    dist.barrier()
    q.put((rank, mine, red))
    dist.destroy_process_group()


def test_two_rank_gloo_max_reduction():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4, 5]
    assert res[0][2] == res[1][2] == [15.0, 2.0]
    assert aggregate_throughput(2000, 2, 10, 15.0) == pytest.approx(2000 * 2 * 10 / 0.015)


def _scatter_worker(rank, world, port, n_streams, q):
    import numpy as np
    import torch
    from botsort_b200.sharding import (frame_slot_ints, gather_streams, pack_frame, pack_result, result_slot_ints,
                                       scatter_streams, unpack_frame, unpack_result)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cap = 16
    rng = np.random.default_rng(5)
    frames = []
    for s in range(n_streams):
        m = int(rng.integers(0, cap + 1))
        frames.append((rng.integers(0, 4000, (m, 4)).astype(np.int32), rng.random(m).astype(np.float32),
                       rng.integers(0, 1000, m)))
    packed = None
    if rank == 0:
        packed = torch.from_numpy(np.stack([pack_frame(b, s, g, cap) for b, s, g in frames]))
    mine_ids = shard_streams(n_streams, world, rank)
    got = scatter_streams(packed, n_streams, frame_slot_ints(cap))
    ok = got.shape[0] == len(mine_ids)
    results = []
    for k, sid in enumerate(mine_ids):
        m, boxes, scores, gt = unpack_frame(got[k], cap)
        b, s, g = frames[sid]
        ok = ok and m == len(b) and np.array_equal(boxes.numpy(), b) and np.array_equal(scores.numpy(), s) \
            and np.array_equal(gt.numpy(), g.astype(np.int32))
        # the "tracker": ids = gt + 1, boxes as float64 shifted by half a pixel
        results.append(pack_result(g.astype(np.int32) + 1, b.astype(np.float64) + 0.5, cap))
    mine = torch.from_numpy(np.stack(results)) if results else torch.empty((0, result_slot_ints(cap)), dtype=torch.int32)
    allres = gather_streams(mine, n_streams, result_slot_ints(cap))
    if rank == 0:
        ok = ok and allres.shape[0] == n_streams
        for sid in range(n_streams):
            ids, tlbr = unpack_result(allres[sid].numpy(), cap)
            b, s, g = frames[sid]
            ok = ok and np.array_equal(ids, g.astype(np.int32) + 1) and np.array_equal(tlbr, b.astype(np.float64) + 0.5)
    else:
        ok = ok and allres is None
    dist.barrier()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams", [8, 5])
def test_two_rank_gloo_scatter_gather(n_streams):
    """BASELINE config 5's data path on CPU: per-stream scatter of detector outputs from rank 0, per-stream gather
    of ids + boxes back (even and ragged partitions)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_scatter_worker, args=(r, 2, port, n_streams, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_numa_local_core_split():
    """Ranks share the cores of the NUMA node their GPU hangs off; unknown topology -> even split (bench.py)."""
    from botsort_b200.sharding import parse_cpulist, split_cores_numa_local
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    node0, node1 = list(range(0, 16)) + list(range(32, 48)), list(range(16, 32)) + list(range(48, 64))
    topo = [node0] * 4 + [node1] * 4
    allowed = list(range(64))
    got = [split_cores_numa_local(topo, allowed, r) for r in range(8)]
    assert all(len(g) == 8 for g in got)
    assert all(set(got[r]) <= set(node0) for r in range(4)) and all(set(got[r]) <= set(node1) for r in range(4, 8))
    assert len(set().union(*map(set, got))) == 64                      # disjoint, everything used
    # no sysfs: the old even split
    got = [split_cores_numa_local([[]] * 4, list(range(32)), r) for r in range(4)]
    assert got == [list(range(8 * r, 8 * r + 8)) for r in range(4)]
    # affinity mask narrower than the node
    assert split_cores_numa_local([node0, node0], list(range(4)), 1) == [2, 3]
