"""The reference's OWN callers, imported unmodified from /root/reference, running against this repository's environments:

    lerobot/lerobot/scripts/eval.py:84-208       rollout()
    lerobot/lerobot/common/envs/utils.py:22-62   preprocess_observation()  (called inside rollout)
    lerobot/lerobot/common/envs/factory.py:22-58 make_env()

The third-party packages those files import and this image lacks (gymnasium, omegaconf, hub internals) are stood in for by
tests/stubs (gymnasium-0.29 semantics); the environments' arithmetic is the CUDA kernel source run through the warp emulator
(tests/emu/emu_capi.py replaces the ctypes layer, nothing else), because the build container has no GPU and the GPU box has no
/root/reference.  Skipped where the reference tree is absent.
"""
import os
import sys

import numpy as np
import pytest

from tests import stubs

pytestmark = pytest.mark.skipif(not stubs.available(), reason="/root/reference is not present on this machine (GPU box)")

HOLD = np.array([0, -0.082, 1.06, 0, -0.953, 0, 1.0] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float32)


@pytest.fixture()
def ref(monkeypatch):
    import torch
    from torch import nn

    stubs.install()
    from av_aloha_b200 import env as avenv
    from tests.emu import emu_capi
    monkeypatch.setattr(avenv, "capi", emu_capi)
    monkeypatch.setattr(avenv, "_MODEL_CACHE", {})
    dropin = os.path.join(os.path.dirname(os.path.abspath(avenv.__file__)), "dropin")
    monkeypatch.syspath_prepend(dropin)
    sys.modules.pop("gym_guided_vision", None)
    import lerobot.scripts.eval as ev                      # the reference's file, as it is
    from lerobot.common.envs.factory import make_env       # idem

    class HoldPolicy(nn.Module):
        def __init__(self):
            super().__init__()
            self.p = nn.Parameter(torch.zeros(1))
            self.calls = 0

        def reset(self):
            self.calls = 0

        def select_action(self, batch):
            assert batch["observation.state"].dtype == torch.float32 and batch["observation.state"].shape[1] == 21
            self.calls += 1
            return torch.as_tensor(HOLD).repeat(batch["observation.state"].shape[0], 1)

    return ev, make_env, avenv, HoldPolicy()


def test_reference_rollout_runs_unmodified_on_the_batched_env(ref):
    ev, make_env, avenv, policy = ref
    import torch
    B, T = 2, 3
    venv = avenv.GuidedVisionVectorEnv("slot_insertion", B, cameras=[], max_episode_steps=T, seed=3)
    out = ev.rollout(venv, policy, seeds=[10, 11], return_observations=True)
    assert policy.calls == T
    assert tuple(out["action"].shape) == (B, T, 21) and tuple(out["reward"].shape) == (B, T)
    assert out["done"].dtype == torch.bool and out["done"][:, -1].all() and not out["done"][:, :-1].any()
    assert not out["success"].any()                                   # holding the home pose inserts nothing
    st = out["observation"]["observation.state"]
    assert tuple(st.shape) == (B, T + 1, 21) and st.dtype == torch.float32
    assert float((st[:, :T] - torch.as_tensor(HOLD)).abs().max()) < 0.05          # arms hold the commanded pose (grippers open = 1)
    assert float((out["reward"] != 0).sum()) == 0
    venv.close()


def test_reference_make_env_serial_envs_equal_the_batched_env(ref):
    """make_env (unchanged) builds SyncVectorEnv([gym.make(id)] * B): B single environments stepped one after the other.  The
    same rollout on this repo's batched environment (reference_rng: object placements drawn from the global np.random in the
    reference's order) must produce the same states bit for bit -- batching is an execution detail, not a different simulator."""
    ev, make_env, avenv, policy = ref
    from omegaconf import DictConfig
    B, T = 2, 2
    cfg = DictConfig({"env": {"name": "guided_vision", "task": "SlotInsertion-3Arms-v0", "episode_length": T,
                              "gym": {"cameras": []}},
                      "eval": {"batch_size": B, "use_async_envs": False}})
    np.random.seed(7)
    env = make_env(cfg)
    assert type(env).__name__ == "SyncVectorEnv" and env.num_envs == B and env.call("_max_episode_steps")[0] == T
    a = ev.rollout(env, policy, seeds=[1, 2], return_observations=True)
    env.close()
    np.random.seed(7)
    venv = avenv.GuidedVisionVectorEnv("slot_insertion", B, cameras=[], max_episode_steps=T, reference_rng=True)
    b = ev.rollout(venv, policy, seeds=[1, 2], return_observations=True)
    venv.close()
    for k in ("action", "reward", "success", "done"):
        assert (a[k] == b[k]).all(), k
    sa, sb = a["observation"]["observation.state"], b["observation"]["observation.state"]
    assert tuple(sa.shape) == (B, T + 1, 21) and (sa == sb).all()
