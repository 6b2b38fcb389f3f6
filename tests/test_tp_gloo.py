"""CPU, world_size 2 over gloo: the tensor-parallel head's exchange step - every rank contributes the
arg-max key of its vocab partition, ONE all_reduce(MAX) over int64 picks the global winner with the
lowest-index tie-break (the reference: topk + all_gather + argmax + gather, arctic_speculator.py:733-744)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from arcticinference_b200 import replicas


def reference_merge(logits_full: torch.Tensor, tp: int, per: int) -> torch.Tensor:
    """arctic_speculator.py:733-744 with the per-rank top-1 taken by argmax: torch.topk does not define
    which of several equal maxima it returns, so on exact ties the reference's TP result is
    implementation-defined; ours is the lowest index = what TP = 1 (torch.argmax, :730-732) returns."""
    B = logits_full.shape[0]
    vals_l, idx_l = [], []
    for r in range(tp):
        logits = logits_full[:, r * per:(r + 1) * per]
        indices = torch.argmax(logits, dim=-1, keepdim=True)
        vals_l.append(torch.gather(logits, -1, indices).to(torch.float64))
        idx_l.append(indices + r * per)
    argidx = torch.argmax(torch.cat(vals_l, -1), -1).reshape(B, -1)
    return torch.gather(torch.cat(idx_l, -1), -1, argidx).flatten()


def make_logits():
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(16, 64, generator=g).to(torch.bfloat16).float()
    logits[0, 5] = logits[0, 40] = 9.0          # exact tie across the two partitions: lower index wins
    logits[1, :] = -1.5                         # every column ties
    logits[2, 33] = 7.0
    logits[3, :] = 0.0
    logits[3, 10] = -0.0                        # signed zeros compare equal
    return logits


def local_keys(logits, rank, per):
    part = logits[:, rank * per:(rank + 1) * per]
    keys = []
    for row in part:
        j = int(torch.argmax(row))              # lowest local index on ties
        keys.append(replicas.pack_argmax_key(float(row[j]), j, rank * per))
    return torch.tensor(keys, dtype=torch.int64)


def test_key_order_single_process():
    logits = make_logits()
    keys = torch.maximum(local_keys(logits, 0, 32), local_keys(logits, 1, 32))
    got = torch.tensor([replicas.unpack_argmax_key(int(k)) for k in keys])
    assert torch.equal(got, reference_merge(logits, 2, 32))
    assert torch.equal(got, torch.argmax(logits, -1))            # = the TP = 1 result
    assert got[0] == 5 and got[1] == 0 and got[2] == 33
    # ordering of the key itself
    k = replicas.pack_argmax_key
    assert k(1.0, 3) > k(0.5, 0) > k(-0.5, 0) > k(-1.0, 0)
    assert k(1.0, 3) > k(1.0, 4) and k(0.0, 1) == k(-0.0, 1)


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    replicas.init_from_env("gloo")
    logits = make_logits()
    keys = replicas.all_reduce_argmax_keys(local_keys(logits, rank, 64 // world))
    q.put((rank, [replicas.unpack_argmax_key(int(k)) for k in keys]))
    replicas.barrier()
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = reference_merge(make_logits(), 2, 32).tolist()
    assert res[0] == exp and res[1] == exp
