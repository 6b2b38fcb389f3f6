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


# ---- tensor-parallel head: host mirror of the arg-max key format (csrc/common.cuh pack_arg,
# gemm_sm100.cu argmax_shard_key_kernel) - used by the CPU tests of the all-reduce plumbing ----------
def pack_argmax_key(logit: float, local_index: int, index_offset: int = 0) -> int:
    """Signed 64-bit key of (logit, global index): larger logit wins, equal logits -> lower index."""
    import struct
    v = logit + 0.0
    u = struct.unpack("<I", struct.pack("<f", v))[0]
    order = (~u & 0xFFFFFFFF) if (u & 0x80000000) else (u | 0x80000000)
    p = (order << 32) | (0xFFFFFFFF - (local_index + index_offset))
    p ^= 0x8000000000000000
    return p - (1 << 64) if p >= (1 << 63) else p


def unpack_argmax_key(key: int) -> int:
    return 0xFFFFFFFF - (key & 0xFFFFFFFF)


def all_reduce_argmax_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """One MAX all-reduce over int64 keys = the reference's topk + all_gather + argmax + gather
    (arctic_speculator.py:733-744)."""
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(keys, op=dist.ReduceOp.MAX, group=group)
    return keys
