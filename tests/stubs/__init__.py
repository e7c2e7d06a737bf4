"""Stand-ins for the third-party packages the reference's eval loop imports and this image does not have -- TEST
INFRASTRUCTURE ONLY.  With them on sys.modules the reference's own files

    lerobot/lerobot/scripts/eval.py              (rollout, eval_policy)
    lerobot/lerobot/common/envs/factory.py       (make_env)
    lerobot/lerobot/common/envs/utils.py         (preprocess_observation)

import UNMODIFIED from /root/reference and run against this repo's environments (tests/test_reference_callers.py).

What is faked and why:
  gymnasium            -- not installable offline.  The surface the reference touches, with gymnasium 0.29.1 semantics (pinned in
                          lerobot/poetry.lock:1204-1205): Env, spaces.Box/Dict, register/make (TimeLimit
                          outermost when max_episode_steps is given), vector.VectorEnv / SyncVectorEnv (reset(seed=list), serial
                          step loop, autoreset with final_observation / final_info object arrays, call()), AsyncVectorEnv = the same.
  omegaconf.DictConfig -- attribute-access dict (make_env reads cfg.env.name, cfg.env.get("gym"), cfg.eval.batch_size ...).
  huggingface_hub.utils._errors -- module path removed in the installed hub 1.x (eval.py:62 imports RepositoryNotFoundError from it).
  lerobot.common.{datasets.factory, logger, policies.factory, utils.io_utils, utils.utils} -- the reference's own modules that
                          eval.py imports at the top but rollout() never calls; they pull in datasets / wandb / hydra / imageio.
                          Replaced by empty shells carrying the imported names (inside_slurm is real: rollout calls it).
Everything rollout() itself executes -- preprocess_observation, get_device_from_parameters, the Policy protocol -- is the
reference's real code.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

REFERENCE_LEROBOT = "/root/reference/lerobot"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_LEROBOT, "lerobot", "scripts"))


# ------------------------------------------------------------------------------------------ gymnasium 0.29 surface
class _Space:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = shape, dtype


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        super().__init__(tuple(shape) if shape is not None else np.shape(low), np.dtype(dtype))
        self.low, self.high = low, high

    def sample(self):
        return np.zeros(self.shape, self.dtype)

    def contains(self, x):
        return np.shape(x) == self.shape


class Dict(_Space, dict):
    def __init__(self, spaces=None, **kw):
        dict.__init__(self, spaces or {}, **kw)
        _Space.__init__(self)
        self.spaces = self


class Env:
    metadata: dict = {}
    spec = None

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class TimeLimit:
    """gymnasium.wrappers.TimeLimit (0.29): truncated = True on the step where elapsed_steps >= max_episode_steps"""

    def __init__(self, env, max_episode_steps):
        self.env, self._max_episode_steps, self._elapsed_steps = env, max_episode_steps, 0

    def __getattr__(self, name):
        if name.startswith("_") and name != "_max_episode_steps":
            raise AttributeError(f"accessing private attribute '{name}' is prohibited")
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            truncated = True
        return obs, reward, terminated, truncated, info

    def render(self):
        return self.env.render()

    def close(self):
        return self.env.close()


_REGISTRY: dict = {}


def register(id, entry_point, kwargs=None, nondeterministic=False, max_episode_steps=None, **_):   # noqa: A002
    _REGISTRY[id] = {"entry_point": entry_point, "kwargs": dict(kwargs or {}), "max_episode_steps": max_episode_steps}


def make(id, max_episode_steps=None, disable_env_checker=None, **kwargs):   # noqa: A002
    spec = _REGISTRY[id]
    ep = spec["entry_point"]
    if isinstance(ep, str):
        mod, name = ep.split(":")
        ep = getattr(importlib.import_module(mod), name)
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    env = ep(**kw)
    steps = max_episode_steps if max_episode_steps is not None else spec["max_episode_steps"]
    return TimeLimit(env, steps) if steps else env


def _stack(obs_list):
    first = obs_list[0]
    if isinstance(first, dict):
        return {k: _stack([o[k] for o in obs_list]) for k in first}
    return np.stack([np.asarray(o) for o in obs_list])


class VectorEnv:
    pass


class SyncVectorEnv(VectorEnv):
    """gymnasium.vector.SyncVectorEnv (0.29): environments stepped one after the other in this process"""

    def __init__(self, env_fns, **_):
        self.envs = [fn() for fn in env_fns]
        self.num_envs = len(self.envs)
        self.metadata = getattr(self.envs[0].unwrapped, "metadata", {})
        self.single_observation_space = getattr(self.envs[0].unwrapped, "observation_space", None)
        self.single_action_space = getattr(self.envs[0].unwrapped, "action_space", None)

    @property
    def unwrapped(self):
        return self

    def reset(self, seed=None, options=None):
        seeds = seed if isinstance(seed, (list, tuple)) else [seed + i if seed is not None else None for i in range(self.num_envs)]
        obs, infos = [], {}
        for env, s in zip(self.envs, seeds):
            o, _info = env.reset(seed=s, options=options)
            obs.append(o)
        return _stack(obs), infos

    def step(self, actions):
        obs, rewards, terms, truncs = [], [], [], []
        final_obs = np.full(self.num_envs, None, dtype=object)
        final_info = np.full(self.num_envs, None, dtype=object)
        is_success = np.zeros(self.num_envs, dtype=object)
        any_final = False
        for i, (env, a) in enumerate(zip(self.envs, actions)):
            o, r, term, trunc, info = env.step(a)
            if term or trunc:
                final_obs[i], final_info[i], any_final = o, info, True
                o, _ = env.reset()
            else:
                is_success[i] = info.get("is_success", False)
            obs.append(o); rewards.append(r); terms.append(term); truncs.append(trunc)
        infos = {}
        if any_final:
            mask = np.array([x is not None for x in final_info])
            infos = {"final_observation": final_obs, "_final_observation": mask, "final_info": final_info, "_final_info": mask.copy()}
        else:
            infos = {"is_success": is_success, "_is_success": np.ones(self.num_envs, bool)}
        return _stack(obs), np.array(rewards, dtype=np.float64), np.array(terms, dtype=bool), np.array(truncs, dtype=bool), infos

    def call(self, name, *args, **kwargs):
        out = []
        for env in self.envs:
            v = getattr(env, name)
            out.append(v(*args, **kwargs) if callable(v) else v)
        return tuple(out)

    def close(self):
        for env in self.envs:
            env.close()


class DictConfig(dict):
    """omegaconf.DictConfig: nested attribute access"""

    def __init__(self, d=None):
        super().__init__({k: DictConfig(v) if isinstance(v, dict) else v for k, v in (d or {}).items()})

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """put the stand-ins on sys.modules and the reference's lerobot package on sys.path; idempotent"""
    if "gymnasium" not in sys.modules or not hasattr(sys.modules["gymnasium"], "__avsim_stub__"):
        spaces = _module("gymnasium.spaces", Box=Box, Dict=Dict)
        vector = _module("gymnasium.vector", VectorEnv=VectorEnv, SyncVectorEnv=SyncVectorEnv, AsyncVectorEnv=SyncVectorEnv)
        wrappers = _module("gymnasium.wrappers", TimeLimit=TimeLimit)
        _module("gymnasium", Env=Env, spaces=spaces, vector=vector, wrappers=wrappers, register=register, make=make,
                __avsim_stub__=True, __version__="0.29.1")
    _module("omegaconf", DictConfig=DictConfig)
    import huggingface_hub.errors as hf_errors
    _module("huggingface_hub.utils._errors", RepositoryNotFoundError=hf_errors.RepositoryNotFoundError)
    if REFERENCE_LEROBOT not in sys.path:
        sys.path.append(REFERENCE_LEROBOT)   # appended: it has its own top-level `tests` package that must not shadow ours
    unused = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("not on the rollout path"))   # noqa: E731
    _module("lerobot.common.datasets.factory", make_dataset=unused)
    _module("lerobot.common.logger", log_output_dir=unused)
    _module("lerobot.common.policies.factory", make_policy=unused)
    _module("lerobot.common.utils.io_utils", write_video=unused)
    _module("lerobot.common.utils.utils", get_safe_torch_device=unused, init_hydra_config=unused, init_logging=unused,
            inside_slurm=lambda: "SLURM_JOB_ID" in os.environ, set_global_seed=unused)
