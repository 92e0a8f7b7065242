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
