"""Scene sharding across the GPUs of one box: scenes are independent, so there is no collective on the data path.
torch.distributed is used only for the timing barrier and the max-over-ranks of the measured times."""
from __future__ import annotations

from typing import List


def shard_scene_indices(scenes_per_gpu: int, rank: int, world: int) -> List[int]:
    """Weak scaling: the job has scenes_per_gpu * world scenes; scene i is rendered by rank i % world."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return [rank + world * k for k in range(scenes_per_gpu)]


def owner_of_scene(scene_idx: int, world: int) -> int:
    return scene_idx % world


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a per-rank time (seconds or ms); identity when torch.distributed is not initialised."""
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
