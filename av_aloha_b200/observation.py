"""Device-resident observation path.

Mirrors, with the same names and return conventions,

* ``preprocess_observation`` (reference lerobot/lerobot/common/envs/utils.py:22-62): environment observation ->
  LeRobot policy inputs: ``observation.images.<cam>`` float32 channel-first in [0, 1], ``observation.state`` float32;
* ``rollout`` (reference lerobot/lerobot/scripts/eval.py:84-215): one batched policy rollout, returning the
  ``action / reward / success / done`` tensors with the same shapes and meaning.

The reference converts on the host, every step: ``torch.from_numpy(img)`` -> ``rearrange b h w c -> b c h w`` ->
``.type(float32)`` -> ``/= 255`` -> ``.to(device)``.  Here the frames never leave HBM: ``avsim_render`` writes uint8
[B, ncam, H, W, 3], ``avsim_pixels_to_float`` (csrc/avsim_obs.cuh) writes float32 [B, ncam, 3, H, W] with the same fp32
division, bit for bit, and the per-camera entries are views of that tensor (ACT stacks them back into exactly this
layout, modeling_act.py:109-110).  Numpy observations (the unchanged gym contract) are accepted as well: they are
uploaded once as uint8 (4x fewer PCIe bytes than the reference's fp32 upload) and converted by the same kernel.
There is no CPU conversion path.
"""
from __future__ import annotations

import numpy as np

from . import capi


def _device_of(observations, device):
    import torch

    if device is not None:
        return torch.device(device)
    for v in (observations.get("pixels"), observations.get("agent_pos")):
        if isinstance(v, dict):
            v = next(iter(v.values()), None)
        if torch.is_tensor(v) and v.is_cuda:
            return v.device
    return torch.device("cuda", torch.cuda.current_device())


def preprocess_observation(observations, cameras=None, device=None):
    """observations: what ``env.step`` / ``env.reset`` returned -- either the gym contract (numpy: ``pixels`` a dict
    ``{cam: uint8 [B, H, W, 3]}`` or one array, ``agent_pos`` float64 [B, nj]) or ``GuidedVisionVectorEnv.observation_device()``
    (``pixels`` one uint8 CUDA tensor [B, ncam, H, W, 3] with ``cameras`` naming axis 1).  Returns CUDA tensors."""
    import torch

    dev = _device_of(observations, device)
    ret = {}
    px = observations.get("pixels")
    if px is not None and not (isinstance(px, dict) and not px):
        if isinstance(px, dict):
            imgs = {f"observation.images.{k}": v for k, v in px.items()}
        elif torch.is_tensor(px) and px.dim() == 5:
            if cameras is None or len(cameras) != px.shape[1]:
                raise ValueError("preprocess_observation: a [B, ncam, H, W, 3] pixel tensor needs `cameras` naming axis 1")
            planes = capi.pixels_to_float(px.contiguous())                      # one launch for all cameras
            imgs = None
            for k, c in enumerate(cameras):
                ret[f"observation.images.{c}"] = planes[:, k]
        else:
            imgs = {"observation.image": px}
        for key, img in (imgs or {}).items():
            if not torch.is_tensor(img):
                img = np.ascontiguousarray(img)
                assert img.dtype == np.uint8, f"expect uint8 images, but instead {img.dtype=}"
                img = torch.from_numpy(img).to(dev, non_blocking=True)
            _, h, w, c = img.shape
            assert c < h and c < w, f"expect channel last images, but instead got {img.shape=}"
            assert img.dtype == torch.uint8, f"expect torch.uint8, but instead {img.dtype=}"
            ret[key] = capi.pixels_to_float(img.to(dev).contiguous())
    if "environment_state" in observations:
        ret["observation.environment_state"] = torch.as_tensor(observations["environment_state"], device=dev).float()
    ret["observation.state"] = torch.as_tensor(observations["agent_pos"], device=dev).float()
    return ret


def rollout(env, policy, seeds=None, return_observations=False, render_callback=None, lazy_render=True):
    """The reference's ``rollout`` on a ``GuidedVisionVectorEnv`` with every tensor resident on the GPU.

    policy: an object with ``reset()`` and ``select_action(batch) -> [B, action_dim]`` (lerobot's Policy protocol).
    lazy_render: policies that execute action chunks (ACT: ``n_action_steps`` = 50, modeling_act.py:123-131) read the
    observation only when their ``_action_queue`` is empty; on the other steps the cameras are not rendered and the previous
    image planes are passed again (they are normalised and discarded by ``select_action``).  Policies without an
    ``_action_queue`` attribute are rendered for on every step.
    Returns {"action" [B, T, A] f32, "reward" [B, T] f64, "success" [B, T] bool, "done" [B, T] bool} (CUDA tensors) and,
    with return_observations, "observation" {key: [B, T+1, ...]}.
    """
    import torch

    policy.reset()
    cams = env.cameras

    def queue_empty():
        q = getattr(policy, "_action_queue", None)
        return q is None or len(q) == 0 or not lazy_render

    observation, _ = env.reset_device(render=True)
    if render_callback is not None:
        render_callback(env)
    all_obs, all_actions, all_rewards, all_successes, all_dones = [], [], [], [], []
    done = torch.zeros(env.num_envs, dtype=torch.bool, device=env._batch.dev)
    max_steps = env.call("_max_episode_steps")[0]
    planes = None
    # every environment of the batch was reset together, so the TimeLimit truncates all of them on step `max_steps`: the
    # reference's `while not np.all(done)` is this loop, without a device->host sync per step
    for _ in range(max_steps):
        if observation["pixels"] is not None or planes is None:
            planes = preprocess_observation(observation, cameras=cams)
        else:                                                   # lazy step: fresh state, previous image planes
            planes = dict(planes)
            planes["observation.state"] = observation["agent_pos"].float()
        if return_observations:
            all_obs.append({k: v.clone() for k, v in planes.items()})
        with torch.inference_mode():
            action = policy.select_action(planes)
        assert action.ndim == 2, "Action dimensions should be (batch, action_dim)"
        observation, reward, terminated, truncated, info = env.step_device(action, render=queue_empty())
        if render_callback is not None:
            render_callback(env)
        success = info["final_success"] if "final_success" in info else torch.zeros_like(done)
        done = terminated | truncated | done
        all_actions.append(action.detach().float())
        all_rewards.append(reward.double())
        all_dones.append(done)
        all_successes.append(success)
    assert bool(done.all()), "rollout: environments still running after max_episode_steps"
    if return_observations:
        if observation["pixels"] is None and cams:
            observation = env.observation_device(render=True)
        all_obs.append({k: v.clone() for k, v in preprocess_observation(observation, cameras=cams).items()})
    ret = {"action": torch.stack(all_actions, dim=1), "reward": torch.stack(all_rewards, dim=1),
           "success": torch.stack(all_successes, dim=1), "done": torch.stack(all_dones, dim=1)}
    if return_observations:
        ret["observation"] = {k: torch.stack([o[k] for o in all_obs], dim=1) for k in all_obs[0]}
    return ret
