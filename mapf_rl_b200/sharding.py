"""Multi-GPU sharding of the environment batch.

Environments are independent (no reference code path couples two `Environment` objects: one env per Ray
actor, worker.py:355-361), so a job of `total` environments is cut into contiguous index ranges, one per
rank / GPU, with NO collective on the data path.  Everything random is keyed by the GLOBAL environment
index (host generator: SeedSequence([seed, index]); device generator: Philox counter = env_offset + slot),
so environment i is the same instance whatever the number of GPUs.  torch.distributed is used only for the
timing barrier and the max-over-ranks reduction of device time in bench.py.
"""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Strong-scaling split of `total` environments: contiguous [lo, hi) of rank `rank`; sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def weak_offset(per_gpu: int, rank: int) -> int:
    """Weak scaling (fixed `per_gpu` environments per rank): global index of the rank's slot 0."""
    return rank * per_gpu


def max_over_ranks(value: float, device=None) -> float:
    """Device time of a multi-GPU job = max over ranks (never wall clock)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_floats(value: float, device=None):
    """One float per rank -> list over ranks (bench.py reports the per-rank spread of the timed region)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def pin_to_cores(local_rank: int, local_world: int) -> list:
    """Give every rank of one box its own contiguous slice of the host cores this process may run on (the host side of the
    host-buffer step -- launch, flag poll -- otherwise migrates between cores that other ranks are spinning on).  Returns
    the cores chosen ([] when the platform has no affinity call or there are fewer cores than ranks)."""
    if local_world <= 1 or not hasattr(os, "sched_getaffinity"):
        return []
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // local_world
        if per < 1:
            return []
        mine = cores[local_rank * per:(local_rank + 1) * per]
        os.sched_setaffinity(0, mine)
        return mine
    except OSError:
        return []
