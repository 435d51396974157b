"""Multi-GPU layout: instances are independent, so the job is a contiguous partition of the
instance index space over ranks (one process per GPU) with NO collective on the step path.
Instance k of the job always uses seed ``base_seed + k``, so results do not depend on the
number of ranks.  The only communication is an off-path all_reduce of a few counters."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def shard_range(num_instances: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[first, last) of the instances owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(num_instances), int(world_size))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def shard_seeds(base_seed: int, num_instances: int, rank: int, world_size: int) -> np.ndarray:
    first, last = shard_range(num_instances, rank, world_size)
    return np.arange(base_seed + first, base_seed + last, dtype=np.uint64)


def aggregate_counters(counters: Dict[str, float], device=None) -> Dict[str, float]:
    """Sum a dict of scalars over all ranks (torch.distributed, NCCL on GPUs / gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([float(counters[k]) for k in keys], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def make_sharded_env(grid_config, num_instances: int, rank: int, world_size: int, device, base_seed=None, **kwargs):
    """This rank's BatchedPogema over its shard of a ``num_instances`` job."""
    from .batched import BatchedPogema
    base = (grid_config.seed or 0) if base_seed is None else base_seed
    seeds = shard_seeds(base, num_instances, rank, world_size)
    kwargs.setdefault("reseed_stride", int(num_instances))  # new seeds never collide across ranks
    return BatchedPogema(grid_config, num_envs=len(seeds), device=device, seeds=seeds, **kwargs)
