"""Config 5 (BASELINE.json): batched policy rollouts, environments sharded across GPUs, device-resident observations.

The shape of lerobot's eval_policy (lerobot/scripts/eval.py:218-410): `n_batches` rollouts of `batch` environments each,
per-episode sum / max reward and success, aggregated at the end.  One process per GPU (torchrun); each rank owns a
contiguous shard of the batch (av_aloha_b200/sharding.py), runs av_aloha_b200/observation.py:rollout on it -- frames
rendered and converted on the device, rendered only when the chunking policy reads them (--lazy) -- and the ranks
exchange ONE all_gather of the per-episode results per rollout.

The reference evaluates an ACT checkpoint (zed_wrist_act.yaml); hub checkpoints and lerobot's own dependencies are not
available offline, so the policy here is a randomly initialised stand-in with ACT's interface and cadence: a small conv
encoder per camera + the joint state -> a chunk of `n_action_steps` = 50 joint targets around the home pose, queued and
popped one per step (modeling_act.py:123-131).  It exercises the loop, not the task: success rates are those of noise.

    python tools/eval_rollout.py [--batch 128] [--rollouts 2] [--steps 300] [--cameras 4] [--no-lazy]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_rollout.py --batch 1024
"""
import argparse
import json
import os
import sys
import time
from collections import deque

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from torch import nn

from av_aloha_b200 import observation, sharding
from av_aloha_b200.env import GuidedVisionVectorEnv

CAMS = ["zed_cam_left", "zed_cam_right", "wrist_cam_left", "wrist_cam_right", "overhead_cam", "worms_eye_cam"]
HOME = [0, -0.082, 1.06, 0, -0.953, 0, 1] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0]


class ChunkPolicy(nn.Module):
    """ACT-shaped stand-in: select_action(batch) pops from a queue that is refilled with n_action_steps actions."""

    def __init__(self, cameras, n_action_steps=50, nj=21):
        super().__init__()
        self.cameras, self.n, self.nj = list(cameras), n_action_steps, nj
        self.enc = nn.Sequential(nn.Conv2d(3, 16, 8, stride=8), nn.ReLU(), nn.Conv2d(16, 32, 4, stride=4), nn.ReLU(),
                                 nn.AdaptiveAvgPool2d(1), nn.Flatten())
        self.head = nn.Linear(32 * max(1, len(self.cameras)) + nj, n_action_steps * nj)
        self.register_buffer("home", torch.tensor(HOME[:nj], dtype=torch.float32))
        self._action_queue = deque([], maxlen=n_action_steps)
        self.forward_calls = 0

    def reset(self):
        self._action_queue.clear()

    @torch.no_grad()
    def select_action(self, batch):
        if len(self._action_queue) == 0:
            self.forward_calls += 1
            state = batch["observation.state"]
            feats = [self.enc(batch[f"observation.images.{c}"]) for c in self.cameras] or [state.new_zeros(len(state), 32)]
            chunk = self.head(torch.cat(feats + [state], dim=1)).view(len(state), self.n, self.nj)
            actions = self.home + 0.05 * torch.tanh(chunk)
            self._action_queue.extend(actions.transpose(0, 1))
        return self._action_queue.popleft()


def run(world, rank, dev, batch_total=1024, task="SlotInsertion", n_cameras=4, steps=300, rollouts=1, lazy=True, height=480, width=640,
        warm_rollout_steps=None):
    """One config-5 measurement on an already initialised process group (or a single process): `rollouts` rollouts of
    `batch_total` environments sharded over the ranks, device-timed (CUDA events, max over ranks).  Returns the result line on
    rank 0, None elsewhere.  warm_rollout_steps: length of the untimed warm-up rollout (default: a full one)."""
    lo, hi = sharding.shard_range(batch_total, rank, world)
    cams = CAMS[:n_cameras]
    env = GuidedVisionVectorEnv(task, hi - lo, cameras=cams, max_episode_steps=steps, device=dev.index, seed=1000 + rank,
                                observation_height=height, observation_width=width)
    torch.manual_seed(0)
    policy = ChunkPolicy(cams).to(dev)
    results = []
    if warm_rollout_steps:                                                  # short warm-up: allocations, queue order
        env._max_episode_steps = int(warm_rollout_steps)
    observation.rollout(env, policy, lazy_render=lazy)
    env._max_episode_steps = steps
    warm_calls = policy.forward_calls
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(rollouts):
        out = observation.rollout(env, policy, lazy_render=lazy)
        done_before = torch.cat([torch.zeros_like(out["done"][:, :1]), out["done"][:, :-1]], dim=1)
        rew = out["reward"].masked_fill(done_before, 0.0)                   # eval.py:283-290: mask steps after the first done
        results.append(sharding.gather_episode_stats(out["success"].any(dim=1), rew.max(dim=1).values.to(torch.int32),
                                                     rew.sum(dim=1).float(), batch_total))
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    wall = time.perf_counter() - t0
    dt = float(ms.item()) * 1e-3
    line = None
    if rank == 0:
        succ, mx, sm = (torch.cat([r[k] for r in results]) for k in range(3))
        line = {"tool": "eval_rollout", "task": task, "n_gpus": world, "batch": batch_total, "rollouts": rollouts,
                "episode_steps": steps, "cameras": cams, "frame": [height, width], "lazy_render": lazy,
                "episodes_per_s": batch_total * rollouts / dt, "env_steps_per_s": batch_total * rollouts * steps / dt,
                "policy_forward_calls_per_rollout": (policy.forward_calls - warm_calls) / rollouts, "device_s": dt, "wall_s": wall,
                "timing": "CUDA events around the rollouts, max over ranks",
                "aggregated": sharding.aggregate(succ, mx, sm), "policy": "random-init ACT-shaped stand-in (chunks of 50)"}
    env.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="SlotInsertion")
    ap.add_argument("--batch", type=int, default=128, help="environments per rollout, all ranks together")
    ap.add_argument("--rollouts", type=int, default=2)
    ap.add_argument("--steps", type=int, default=300, help="episode length (sim_slot_insertion_3arms.yaml:17)")
    ap.add_argument("--cameras", type=int, default=4)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--no-lazy", action="store_true", dest="no_lazy")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = run(world, rank, dev, args.batch, args.task, args.cameras, args.steps, args.rollouts, not args.no_lazy, args.height, args.width)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
