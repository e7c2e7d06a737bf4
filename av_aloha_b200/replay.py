"""Batched dataset replay on the GPU.

Two reference scripts walk recorded episodes one frame / one step at a time through a single MuJoCo env:

* ``replay_sim_episode.py`` (reference gym_guided_vision/scripts/replay_sim_episode.py:47-89 and
  data_collection_scripts/replay_sim_episode.py:47-93): ``for qpos in all_qpos: env.set_qpos(qpos); obs = env.get_obs()``
  -- re-renders a recorded episode for a new camera set.  Frames are independent, so here every frame of the episode is
  one *environment* of a lockstep batch: one ``avsim_set(QPOS)`` + ``avsim_forward`` + ``avsim_render`` per chunk.
* ``check_dataset_reward.py`` (reference gym_guided_vision/scripts/check_dataset_reward.py:28-55):
  ``env.set_qpos(all_qpos[0]); for a in actions: env.step_action(a); reward.append(env.get_reward())`` and the episode
  passes when ``max(reward) == env.max_reward``.  Episodes are independent, so here every *episode* is one environment and
  the whole dataset advances one recorded action per launch.

Inputs are plain arrays (the reference reads them from HDF5 ``/observations/all_qpos`` and ``/action``,
record_sim_episodes.py:161-210; h5py is not part of this package).  There is no CPU path: both functions run the CUDA
kernels through the C-ABI and raise without a device.
"""
from __future__ import annotations

import numpy as np

from . import capi, model_io
from .env import SIM_PHYSICS_ENV_STEP_RATIO, TASK_OF, _camera_ids, _check_cameras, _model


def rerender_episode(task: str, all_qpos, cameras, num_arms: int = 3, height: int = 480, width: int = 640,
                     device: int = 0, chunk: int = 256, to_numpy: bool = True):
    """Frames of every configured camera for every recorded configuration.

    all_qpos: [T, nq] (the model's full qpos, robot + task objects, as stored in ``/observations/all_qpos``).
    Returns {camera: uint8 [T, H, W, 3]} (numpy, or CUDA tensors with ``to_numpy=False``), i.e. the
    ``/observations/images/<cam>`` datasets the reference script writes.
    """
    import torch

    task = TASK_OF.get(task, task)
    cameras = list(cameras)
    _check_cameras(cameras)
    model = _model(task, num_arms, device)
    q = np.ascontiguousarray(all_qpos, np.float32)
    if q.ndim != 2 or q.shape[1] != model.nq:
        raise ValueError(f"rerender_episode: all_qpos must be [T, {model.nq}], got {q.shape}")
    T = q.shape[0]
    out = {c: (np.empty((T, height, width, 3), np.uint8) if to_numpy else
               torch.empty((T, height, width, 3), dtype=torch.uint8, device=torch.device("cuda", device))) for c in cameras}
    if T == 0 or not cameras:
        return out
    cam_ids = _camera_ids(task, num_arms, cameras)
    n = min(int(chunk), T)
    batch = capi.Batch(model, n, seed=0)
    try:
        batch.reset()
        frames = None
        for t0 in range(0, T, n):
            rows = q[t0:t0 + n]
            m = rows.shape[0]
            if m < n:                                    # ragged tail: pad with the last row, drop the padding below
                rows = np.concatenate([rows, np.repeat(rows[-1:], n - m, axis=0)])
            batch.set(capi.QPOS, rows)
            batch.forward()                              # set_qpos = write qpos + physics.forward() (reference env.py:251-253)
            frames = batch.render(cam_ids, height, width, out=frames)
            for k, c in enumerate(cameras):
                if to_numpy:
                    out[c][t0:t0 + m] = frames[:m, k].cpu().numpy()
                else:
                    out[c][t0:t0 + m] = frames[:m, k]
    finally:
        batch.close()
    return out


def audit_rewards(task: str, first_qpos, actions, num_arms: int = 3, device: int = 0, solver: str = "newton",
                  solver_iterations: int = 8, warmstart: int = 2, lengths=None):
    """Replays recorded actions for all episodes at once and reports which reach the maximum reward.

    first_qpos: [E, nq] (``all_qpos[0]`` of every episode); actions: [E, T, 14|21] float (``/action``), padded to a
    common T; lengths: optional [E] valid lengths (steps past an episode's length repeat its last action and are
    ignored in the maximum).  Returns a dict with ``max_reward_reached`` (bool [E]), ``episode_max`` (int [E]),
    ``rewards`` (int [E, T]) and ``not_max_reward_episodes`` (indices, what the reference script prints).
    """
    import torch

    task = TASK_OF.get(task, task)
    model = _model(task, num_arms, device)
    q0 = np.ascontiguousarray(first_qpos, np.float32)
    a = np.ascontiguousarray(actions, np.float32)
    if q0.ndim != 2 or q0.shape[1] != model.nq:
        raise ValueError(f"audit_rewards: first_qpos must be [E, {model.nq}], got {q0.shape}")
    E = q0.shape[0]
    if a.ndim != 3 or a.shape[0] != E or a.shape[2] < model.njoints:
        raise ValueError(f"audit_rewards: actions must be [E, T, >={model.njoints}], got {a.shape}")
    a = np.ascontiguousarray(a[:, :, :model.njoints])    # 2-arm envs read the first 14 entries (reference env.py:204-215)
    T = a.shape[1]
    lengths = np.full(E, T, np.int64) if lengths is None else np.asarray(lengths, np.int64)
    if lengths.shape != (E,) or (lengths < 0).any() or (lengths > T).any():
        raise ValueError("audit_rewards: lengths must be [E] with 0 <= length <= T")
    rewards = np.zeros((E, T), np.int32)
    if E == 0 or T == 0:
        return {"max_reward_reached": np.zeros(E, bool), "episode_max": np.zeros(E, np.int32), "rewards": rewards,
                "not_max_reward_episodes": list(range(E))}
    dev = torch.device("cuda", device)
    batch = capi.Batch(model, E, seed=0)
    try:
        batch.set_solver(solver)
        if solver == "pgs":
            batch.set_options(solver_iters=solver_iterations)
            batch.set_warmstart(warmstart)
        batch.reset()                                    # env.reset(options={}) ...
        batch.set(capi.QPOS, q0)                         # ... then env.unwrapped.set_qpos(all_qpos[0])
        batch.forward()
        for e in range(E):                               # hold the last valid action past an episode's end
            if 0 < lengths[e] < T:
                a[e, lengths[e]:] = a[e, lengths[e] - 1]
        acts = torch.as_tensor(a, device=dev).transpose(0, 1).contiguous()    # [T, E, nj] resident in HBM
        rew = torch.empty((T, E), dtype=torch.int32, device=dev)
        for t in range(T):
            batch.step(acts[t], SIM_PHYSICS_ENV_STEP_RATIO)
            batch.get(capi.REWARD, out=rew[t])
        rewards = rew.cpu().numpy().T.copy()
    finally:
        batch.close()
    valid = np.arange(T)[None, :] < lengths[:, None]
    episode_max = np.where(valid, rewards, 0).max(axis=1).astype(np.int32)
    ok = (episode_max == model.max_reward) & (lengths > 0)
    return {"max_reward_reached": ok, "episode_max": episode_max, "rewards": rewards,
            "not_max_reward_episodes": [int(i) for i in np.nonzero(~ok)[0]]}
