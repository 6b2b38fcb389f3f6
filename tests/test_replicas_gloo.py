"""CPU, world_size 2 over gloo: the N>1 path of the hot path is 'independent replicas, no
data-path collective' (SURVEY 8e).  What is distributed is only the plumbing bench.py uses:
batch sharding, the barrier and the max-time / sum-units reduction."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arcticinference_b200 import replicas


def test_shard_rows_partitions_exactly():
    for total in (0, 1, 7, 64, 256, 513):
        for world in (1, 2, 3, 4, 8):
            spans = [replicas.shard_rows(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for start, count in spans:
                assert start == pos
                pos += count
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        replicas.shard_rows(4, 2, 2)


def test_aggregate_single_process():
    assert replicas.aggregate(2.5, 192) == (2.5, 192.0)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, lr, w = replicas.init_from_env("gloo")
    assert (r, w) == (rank, world)
    start, count = replicas.shard_rows(129, w, r)
    # each replica "processes" its own rows; rank 1 is slower
    replicas.barrier()
    ms, units = replicas.aggregate(10.0 * (rank + 1), count * 3)
    q.put((rank, start, count, ms, units))
    replicas.barrier()
    dist.destroy_process_group()


def test_two_replicas_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, c0, ms0, u0), (r1, s1, c1, ms1, u1) = res
    assert (s0, c0, s1, c1) == (0, 65, 65, 64)
    assert ms0 == ms1 == 20.0            # max over ranks
    assert u0 == u1 == 129 * 3           # sum over ranks: whole-job units
