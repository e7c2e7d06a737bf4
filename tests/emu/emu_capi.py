"""A stand-in for av_aloha_b200.capi whose Model / Batch run the CUDA kernel SOURCE through the warp emulator (emu.py) -- TEST
INFRASTRUCTURE ONLY.  It lets the host-side environment classes (av_aloha_b200/env.py) run in the GPU-less build container, so
that the reference's unmodified rollout loop can be executed against them (tests/test_reference_callers.py).  Rendering is not
emulated.  Nothing in the product imports this."""
from __future__ import annotations

import numpy as np
import torch

from av_aloha_b200 import model_io
from av_aloha_b200.capi import (AGENT_POS, CONTACTS, CTRL, LATCH, NCON, QACC, QPOS, QVEL, REWARD, SOLVER_STAT,  # noqa: F401
                                STATUS, SUCCESS, WARMSTART)

from .emu import EmuBatch

_NAMES = {QPOS: "qpos", QVEL: "qvel", CTRL: "ctrl", WARMSTART: "warm", AGENT_POS: "agent_pos", REWARD: "reward",
          NCON: "ncon", STATUS: "status", LATCH: "latch", QACC: "qacc", SOLVER_STAT: "nw_stat"}


class Model:
    def __init__(self, avm_path, device=0):
        self.avm_path = avm_path
        self._t = model_io.load_avm(avm_path)
        self.num_arms = int(self._t["num_arms"][0])
        self.njoints = 21 if self.num_arms == 3 else 14
        self.max_reward = int(self._t["max_reward"][0])
        self.nq, self.nv, self.nu = len(self._t["qpos0"]), len(self._t["dof_damping"]), len(self._t["act_kp"])

    def table(self, name):
        return self._t[name]


class Batch:
    def __init__(self, model, num_envs, seed=0, stream=None):
        self.model, self.num_envs = model, int(num_envs)
        self.eb = EmuBatch(model.avm_path, self.num_envs)
        self.dev = torch.device("cpu")
        self.launch_count = 0
        self.eb.reset()

    def set_solver(self, solver="newton", max_iter=0, ls_iter=0, tol=0.0):
        self.eb.set_solver(solver, max_iter or 30, ls_iter or 20, tol or 3e-7)

    def set_options(self, solver_iters=20, noslip_iters=-1, multiccd=-1):
        self.eb.set_options(solver_iters, noslip_iters, multiccd)

    def set_warmstart(self, mode):
        self.eb.set_warmstart(mode)

    def reset(self, mask=None, free_pos=None):
        if isinstance(mask, torch.Tensor):
            mask = mask.cpu().numpy()
        if isinstance(free_pos, torch.Tensor):
            free_pos = free_pos.cpu().numpy()
        self.eb.reset(free_pos=free_pos, mask=mask)

    def step_host(self, action_np, nsubsteps=20, agent_pos_out=None, reward_out=None, status_out=None):
        self.eb.step(np.ascontiguousarray(action_np, np.float32), nsubsteps)
        if agent_pos_out is not None:
            agent_pos_out[:] = self.eb.agent_pos
        if reward_out is not None:
            reward_out[:] = self.eb.reward
        if status_out is not None:
            status_out[:] = self.eb.status
        return agent_pos_out, reward_out

    def step(self, action, nsubsteps=20):
        self.eb.step(np.ascontiguousarray(action.cpu().numpy(), np.float32), nsubsteps)

    def forward(self):
        self.eb.forward()

    def get(self, field, out=None):
        if field == SUCCESS:
            return torch.as_tensor((self.eb.reward == self.model.max_reward).astype(np.int32))
        return torch.as_tensor(np.array(getattr(self.eb, _NAMES[field])))

    def set(self, field, value):
        v = value.cpu().numpy() if isinstance(value, torch.Tensor) else np.asarray(value)
        getattr(self.eb, _NAMES[field])[:] = v

    def render(self, *a, **k):
        raise RuntimeError("the emulation harness has no renderer: construct the environment with cameras=[]")

    def close(self):
        self.eb = None
