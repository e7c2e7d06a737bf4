"""Drop-in `gym_guided_vision` package: put `<repo>/av_aloha_b200/dropin` on PYTHONPATH ahead of the reference's package and
`import gym_guided_vision` -- what lerobot's make_env does (lerobot/common/envs/factory.py:33-36) -- registers the reference's ten
environment ids (reference gym_guided_vision/__init__.py:4-101: same ids, same kwargs, nondeterministic=True) with entry points
in this repository's CUDA-backed environment classes.  Needs gymnasium, like the package it replaces."""
import gymnasium as gym

from av_aloha_b200.env import ENVS, _TASK_CLASSES

for _e in ENVS:
    gym.register(id=_e["id"], entry_point=f"av_aloha_b200.env:{_TASK_CLASSES[_e['task']].__name__}", kwargs=dict(_e["kwargs"]),
                 nondeterministic=True)
