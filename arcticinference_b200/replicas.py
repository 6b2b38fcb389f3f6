"""Data-parallel replicas: both halves of the hot path shard by request batch with
NO collective on the data path - each GPU owns its own weights copy, its own paged KV cache and
its own slice of the batch.  ``torch.distributed`` is used only for the launch plumbing of
bench.py / tests: a barrier around the timed region and the max-over-ranks / sum-over-ranks
reduction of the measurements."""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_rows(total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of ``total`` request rows over ``world_size`` replicas; the first
    ``total % world_size`` ranks take one extra row.  Returns (start, count)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(total, world_size)
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when absent."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: str, device: torch.device | None = None) -> Tuple[int, int, int]:
    rank, local_rank, world = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if (device is not None and backend == "nccl") else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def barrier() -> None:
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def aggregate(elapsed_ms: float, units: float, device: torch.device | str = "cpu") -> Tuple[float, float]:
    """(max over ranks of elapsed_ms, sum over ranks of units).  Whole-job throughput is
    ``sum_units / max_ms`` - replicas run concurrently, the slowest one bounds the job."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(elapsed_ms), float(units)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
