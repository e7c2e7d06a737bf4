"""Environment sharding across ranks (one process per GPU) and the one collective of the path.

Environments are independent (no cross-environment term anywhere in the reference's env.py), so rank r simply owns the
contiguous slice [r*B/G, (r+1)*B/G) and steps it on its own GPU with no data-path communication.  The only exchange
is at episode end: an all_gather of per-environment `success`, `max_reward` and `sum_reward` -- what lerobot's
eval_policy aggregates (lerobot/scripts/eval.py:295-308, 394-401) -- 9 bytes per environment over NCCL (NVLink /
NVSwitch; latency-bound).  The same code runs over gloo on CPU tensors, which is how the tests cover it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(num_envs: int, rank: int, world: int):
    """Contiguous slice of environments owned by `rank` (the first num_envs % world ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(num_envs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_episode_stats(success: torch.Tensor, max_reward: torch.Tensor, sum_reward: torch.Tensor, num_envs: int,
                         group=None):
    """all_gather the per-environment episode results of every rank's shard, in global environment order.

    success: bool/uint8 [b], max_reward: int32 [b], sum_reward: float32 [b] (b = this rank's shard size; shards may
    differ by one).  Returns (success uint8 [num_envs], max_reward int32 [num_envs], sum_reward float32 [num_envs]).
    """
    if not dist.is_available() or not dist.is_initialized():
        return success.to(torch.uint8), max_reward.to(torch.int32), sum_reward.to(torch.float32)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(num_envs, rank, world)
    if success.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {success.shape[0]} environments, its shard is {hi - lo}")
    width = -(-num_envs // world)                       # pad shards to a common length for all_gather
    pack = torch.zeros((width, 3), dtype=torch.float32, device=success.device)
    pack[: hi - lo, 0] = success.to(torch.float32)
    pack[: hi - lo, 1] = max_reward.to(torch.float32)
    pack[: hi - lo, 2] = sum_reward.to(torch.float32)
    out = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(out, pack, group=group)
    rows = []
    for r in range(world):
        a, b = shard_range(num_envs, r, world)
        rows.append(out[r][: b - a])
    full = torch.cat(rows)
    return full[:, 0].to(torch.uint8), full[:, 1].to(torch.int32), full[:, 2]


def aggregate(success, max_reward, sum_reward):
    """The aggregate fields of lerobot's eval JSON (eval.py:394-401) from gathered per-episode results."""
    return {"avg_sum_reward": float(sum_reward.float().mean()), "avg_max_reward": float(max_reward.float().mean()),
            "pc_success": float(success.float().mean() * 100.0)}
