"""Host-side mirror of gym_guided_vision's environment API on top of the CUDA batch (C-ABI, csrc/).

Same names, argument meaning and return conventions as the reference (gym_guided_vision/gym_guided_vision/env.py):

  GuidedVisionEnv.reset / step / get_obs / get_reward / render / set_qpos / step_action / hide_middle_arm /
  show_middle_arm / close, the five task subclasses, `make_sim_env`, `ENVS` + `register` (same ten ids and kwargs,
  gym_guided_vision/__init__.py:4-101), plus `GuidedVisionVectorEnv`, a batched environment with gymnasium 0.29
  `SyncVectorEnv` semantics (what lerobot's `rollout` / `eval_policy` are written against: eval.py:126-176).

Every environment instance owns rows of one `capi.Batch`; all arithmetic (ctrl write, 20 substeps, reward, agent_pos)
runs in the hand-written kernels.  There is no CPU path: constructing an environment without a CUDA device raises.
gymnasium is optional (not installed in the build container): when importable the classes subclass gym.Env and use
gym spaces, otherwise small stand-ins with the same attributes are used.
"""
from __future__ import annotations

import numpy as np

from . import capi, model_io

try:  # pragma: no cover - gymnasium is absent in the build container
    import gymnasium as gym
    from gymnasium import spaces
    _EnvBase = gym.Env
except Exception:  # noqa: BLE001
    gym = None

    class _Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

        def sample(self):
            return np.zeros(self.shape, self.dtype)

        def contains(self, x):
            return np.shape(x) == self.shape

    class _Dict(dict):
        def __init__(self, d):
            super().__init__(d)
            self.spaces = self

    class spaces:  # noqa: N801 - mirrors the gymnasium namespace
        Box, Dict = _Box, _Dict

    class _EnvBase:
        metadata: dict = {}

        def reset(self, seed=None, options=None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

        @property
        def unwrapped(self):
            return self


# ---- reference constants.py:20-28
SIM_DT = 0.04
SIM_PHYSICS_DT = 0.002
SIM_PHYSICS_ENV_STEP_RATIO = int(SIM_DT / SIM_PHYSICS_DT)
LEFT_ARM_POSE = [0, -0.082, 1.06, 0, -0.953, 0, 0.02239]
RIGHT_ARM_POSE = [0, -0.082, 1.06, 0, -0.953, 0, 0.02239]
MIDDLE_ARM_POSE = [0, -0.8, 0.8, 0, 0.5, 0, 0]
CAMERAS = ["zed_cam_left", "zed_cam_right", "wrist_cam_left", "wrist_cam_right", "overhead_cam", "worms_eye_cam"]
# the 2-arm ids carry no zed cameras (they sit on the middle arm, which the 2-arm model parks out of view):
# reference gym_guided_vision/__init__.py:16,32,48,64,81
CAMERAS_2ARMS = ["overhead_cam", "worms_eye_cam", "wrist_cam_left", "wrist_cam_right"]
RENDER_CAMERA = "overhead_cam"

TASK_OF = {"InsertPeg": "insert_peg", "SlotInsertion": "slot_insertion", "SewNeedle": "sew_needle",
           "TubeTransfer": "tube_transfer", "HookPackage": "hook_package"}

# ---- object placement of the task resets, as (joint, lo[3], hi[3]) draws in the reference's np.random ORDER, dead draws
# included (joint None), shared draws as a string alias.  reference env.py:478-493, 517-533, 608-624, 709-723, 796-808.
_RESET_DRAWS = {
    "insert_peg": [("peg_joint", [0.1, -0.1, 0.01], [0.2, 0.1, 0.01]), ("hole_joint", [-0.1, -0.1, 0.021], [-0.2, 0.1, 0.021])],
    "slot_insertion": [("slot_joint", [-0.05, 0.1, 0.0], [0.05, 0.15, 0.0]), (None, [-0.05, 0.1, 0.0], [0.05, 0.15, 0.0]),
                       ("stick_joint", [-0.08, -0.1, 0.0], [0.08, 0.0, 0.0])],
    "sew_needle": [("needle_joint", [0.15, -0.025, 0.0], [0.2, 0.1, 0.0]), (None, [0.15, -0.025, 0.0], [0.2, 0.1, 0.0]),
                   ("wall_joint", [-0.025, -0.025, 0.0], [0.025, 0.1, 0.0])],
    "tube_transfer": [("ball_joint", [0.05, -0.05, 0.0], [0.1, 0.05, 0.0]), ("tube1_joint", "ball_joint", None),
                      ("tube2_joint", [-0.1, -0.05, 0.0], [-0.05, 0.05, 0.0])],
    "hook_package": [("hook_joint", [-0.1, 0.3, 0.2], [0.1, 0.3, 0.3]), ("package_joint", [-0.1, 0.0, 0.0], [0.1, 0.15, 0.0])],
}


def reference_reset_draws(task: str, free_joint_names, rng=None) -> np.ndarray:
    """Object positions [nfree, 3] of one task reset, consuming `rng` (default: the GLOBAL np.random, like the reference)
    in exactly the reference's order, dead draws included."""
    rng = np.random if rng is None else rng
    pos = {}
    for joint, lo, hi in _RESET_DRAWS[task]:
        if isinstance(lo, str):
            pos[joint] = pos[lo]
            continue
        lo_a, hi_a = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        # == legacy np.random.uniform(lo, hi) bit for bit (lo + (hi - lo) * u), also for the reference's reversed range
        # x in [-0.1, -0.2] (env.py:485), which numpy's new Generator.uniform refuses
        u = rng.random_sample(3) if hasattr(rng, "random_sample") else rng.random(3)
        p = lo_a + (hi_a - lo_a) * u
        if joint is not None:
            pos[joint] = p
    return np.stack([pos[j] for j in free_joint_names])


_MODEL_CACHE: dict = {}


def _model(task, num_arms, device):
    key = (task, num_arms, device)
    if key not in _MODEL_CACHE:
        _MODEL_CACHE[key] = capi.Model(model_io.model_path(task, num_arms), device)
    return _MODEL_CACHE[key]


def _configure_solver(batch, solver, solver_iterations, warmstart):
    """One default for every environment class: "newton" = Newton on the primal run to its tolerance, the solver the reference
    uses (assets/aloha_sim.xml:4-6 leaves MuJoCo's default).  "pgs" = the fixed-cost approximate mode: `solver_iterations`
    block Gauss-Seidel sweeps on the dual with warm start `warmstart` (1: previous qacc, 2: per-constraint force cache)."""
    batch.set_solver(solver)
    if solver == "pgs":
        batch.set_options(solver_iters=solver_iterations)
        batch.set_warmstart(warmstart)


def _check_cameras(cameras):
    assert all(c in CAMERAS for c in cameras), f"Invalid camera names: {cameras}"


def _camera_ids(task, num_arms, cameras):
    names = model_io.load_names(task, num_arms)["camera"]
    return [names.index(c) for c in cameras]


class GuidedVisionEnv(_EnvBase):
    """One environment (reference env.py:32-406).  `task` replaces the reference's `xml` path."""

    metadata = {"render_modes": ["rgb_array"], "render_fps": 1 / SIM_DT}
    task = None
    max_reward = 0

    def __init__(self, task: str | None = None, num_arms: int = 3, cameras=(), observation_height: int = 480,
                 observation_width: int = 640, device: int = 0, solver: str = "newton", solver_iterations: int = 50, seed: int = 0):
        assert num_arms in [2, 3], f"Invalid number of arms: {num_arms}"
        self.task = task or self.task
        self.cameras = list(cameras)
        _check_cameras(self.cameras)
        self.num_arms, self.num_joints = num_arms, 14 if num_arms == 2 else 21
        self.observation_height, self.observation_width = observation_height, observation_width
        self._model = _model(self.task, num_arms, device)
        self._model_arms, self._device, self._seed, self._solver_cfg = num_arms, device, seed, (solver, solver_iterations, 1)
        self._batch = capi.Batch(self._model, 1, seed=seed)
        _configure_solver(self._batch, solver, solver_iterations, 1)
        self.max_reward = self._model.max_reward
        self._free_joints = model_io.load_names(self.task, num_arms)["free_joint"]
        self.observation_space = spaces.Dict({
            "pixels": spaces.Dict({c: spaces.Box(low=0, high=255, shape=(observation_height, observation_width, 3),
                                                 dtype=np.uint8) for c in self.cameras}),
            "agent_pos": spaces.Box(low=-np.inf, high=np.inf, shape=(self.num_joints,), dtype=np.float64)})
        self.action_space = spaces.Box(low=-np.inf, high=np.inf, shape=(self.num_joints,), dtype=np.float32)
        import torch
        self.torch = torch
        self._agent = np.zeros((1, self.num_joints), np.float32)
        self._reward = np.zeros((1,), np.int32)
        self._cam_ids = _camera_ids(self.task, num_arms, self.cameras)
        self._render_cam = _camera_ids(self.task, num_arms, [RENDER_CAMERA])

    def _pixels(self):
        if not self.cameras:
            return {}
        img = self._batch.render(self._cam_ids, self.observation_height, self.observation_width).cpu().numpy()
        return {c: img[0, k] for k, c in enumerate(self.cameras)}

    # -- observation / reward (reference env.py:168-193)
    def _agent_pos(self):
        agent = self._batch.get(capi.AGENT_POS).cpu().numpy()[0].astype(np.float64)
        if self._model_arms == self.num_arms:
            return agent
        # middle arm hidden on a 3-arm environment: the swapped-in 2-arm tables publish 14 joints; the middle arm's seven come
        # straight from qpos (reference env.py:169-178 reads them through the same named joints whatever the base pose is)
        qadr = self._model.table("obs_qadr")
        q = self._batch.get(capi.QPOS).cpu().numpy()[0].astype(np.float64)
        return np.concatenate([agent[:14], q[qadr[14:21]]])

    def get_obs(self):
        return {"pixels": self._pixels(), "agent_pos": self._agent_pos()}

    def get_reward(self):
        return int(self._batch.get(capi.REWARD).cpu().numpy()[0])

    def render(self):
        """225 x 300 frame of the overhead camera (reference env.py:195-200)"""
        return self._batch.render(self._render_cam, 225, 300).cpu().numpy()[0, 0]

    # -- viewer (reference env.py:373-392 opens mujoco.viewer.launch_passive, an interactive GL window, and syncs it).  There is
    # no window system behind a batched GPU simulator; the two calls keep their names and meaning as far as a headless process
    # can: create_viewer registers the key callback (never invoked: there is no keyboard), render_viewer refreshes `viewer_frame`,
    # a 480 x 640 overhead frame a caller may show with whatever it has (cv2.imshow, a notebook), and returns it.
    def create_viewer(self, key_callback=None) -> None:
        self._viewer_key_callback = key_callback
        self.viewer_frame = None

    def render_viewer(self):
        if not hasattr(self, "viewer_frame"):
            self.create_viewer()
        self.viewer_frame = self._batch.render(self._render_cam, 480, 640).cpu().numpy()[0, 0]
        return self.viewer_frame

    # -- stepping (reference env.py:203-226, 255-269)
    def step_action(self, action):
        a = np.asarray(action, np.float32).reshape(1, self.num_joints)
        nj = self._model.njoints
        if nj != self.num_joints:      # middle arm hidden: its ctrl is still written (reference env.py:213-215), it just is not in view
            ctrl = self._batch.get(capi.CTRL)
            ctrl[0, 14:21] = self.torch.as_tensor(a[0, 14:21], device=ctrl.device)
            self._batch.set(capi.CTRL, ctrl)
        agent = np.zeros((1, nj), np.float32)
        self._batch.step_host(np.ascontiguousarray(a[:, :nj]), SIM_PHYSICS_ENV_STEP_RATIO, agent, self._reward)
        self._agent[0, :nj] = agent[0]

    def step(self, action):
        self.step_action(action)
        observation = {"pixels": self._pixels(), "agent_pos": self._agent_pos()}
        reward = int(self._reward[0])
        return observation, reward, False, False, {"is_success": reward == self.max_reward}

    # -- reset: home pose + task object placement drawn from the GLOBAL np.random in the reference's order
    def reset(self, seed=None, options=None):
        super().reset(seed=seed, options=options)
        fp = reference_reset_draws(self.task, self._free_joints)
        self._batch.reset(free_pos=fp[None])
        return self.get_obs(), {"is_success": False}

    def set_qpos(self, qpos):
        self._batch.set(capi.QPOS, np.asarray(qpos, np.float32).reshape(1, -1))
        self._batch.forward()

    # reference env.py:394-398: move middle_base_link to (0, -2.4, -0.4) / back to (0, -0.513, 0.02).  The model tables are
    # compiled per arm count with exactly that edit (mjcf_compile.py), so the call swaps the compiled model under the same state:
    # joint positions, velocities and ctrl carry over (the two models have identical dof / actuator layouts, reference env.py:60-62)
    def hide_middle_arm(self):
        self._swap_model(2)

    def show_middle_arm(self):
        self._swap_model(3)

    def _swap_model(self, arms):
        if arms == self._model_arms:
            return
        state = {f: self._batch.get(f).clone() for f in (capi.QPOS, capi.QVEL, capi.CTRL, capi.WARMSTART, capi.LATCH)}
        self._batch.close()
        self._model = _model(self.task, arms, self._device)
        self._model_arms = arms
        self._batch = capi.Batch(self._model, 1, seed=self._seed)
        _configure_solver(self._batch, *self._solver_cfg)
        for f, v in state.items():
            self._batch.set(f, v)
        self._batch.forward()
        self._cam_ids = _camera_ids(self.task, arms, self.cameras)
        self._render_cam = _camera_ids(self.task, arms, [RENDER_CAMERA])

    def close(self):
        if getattr(self, "_batch", None) is not None:
            self._batch.close()
            self._batch = None


class InsertPegEnv(GuidedVisionEnv):
    task = "insert_peg"


class SlotInsertionEnv(GuidedVisionEnv):
    task = "slot_insertion"


class SewNeedleEnv(GuidedVisionEnv):
    task = "sew_needle"


class TubeTransferEnv(GuidedVisionEnv):
    task = "tube_transfer"


class HookPackageEnv(GuidedVisionEnv):
    task = "hook_package"


_TASK_CLASSES = {"insert_peg": InsertPegEnv, "slot_insertion": SlotInsertionEnv, "sew_needle": SewNeedleEnv,
                 "tube_transfer": TubeTransferEnv, "hook_package": HookPackageEnv}


def make_sim_env(task_name, **kwargs):
    """reference env.py:18-30"""
    for key, cls in _TASK_CLASSES.items():
        if f"sim_{key}" in task_name:
            return cls(**kwargs)
    raise NotImplementedError


# ---- registry (reference gym_guided_vision/__init__.py:4-101): same ids, same kwargs
ENVS = []
for _name in TASK_OF:
    for _arms in (2, 3):
        ENVS.append({"id": f"gym_guided_vision/{_name}-{_arms}Arms-v0", "task": TASK_OF[_name],
                     "kwargs": {"num_arms": _arms, "cameras": list(CAMERAS if _arms == 3 else CAMERAS_2ARMS),
                                "observation_height": 480, "observation_width": 640}})


def make(env_id: str, max_episode_steps: int | None = None, **kwargs):
    """`gym.make` for the ten ids without needing gymnasium; kwargs override the registered defaults."""
    spec = next((e for e in ENVS if e["id"] == env_id), None)
    if spec is None:
        raise KeyError(f"unknown environment id {env_id!r}")
    kw = dict(spec["kwargs"])
    kw.update(kwargs)
    env = _TASK_CLASSES[spec["task"]](**kw)
    env._max_episode_steps = max_episode_steps
    return env


def register():
    """Register the ten ids with gymnasium (when it is installed) under the reference's names, pointing at this backend."""
    if gym is None:
        raise RuntimeError("gymnasium is not installed")
    for e in ENVS:
        cls = _TASK_CLASSES[e["task"]].__name__
        gym.register(id=e["id"], entry_point=f"av_aloha_b200.env:{cls}", kwargs=e["kwargs"], nondeterministic=True)


class GuidedVisionVectorEnv:
    """B environments in lockstep behind gymnasium-0.29 `SyncVectorEnv` semantics.

    What lerobot's rollout uses (lerobot/scripts/eval.py:126-176, 259-263, 321): `num_envs`, `reset(seed=list|int|None)`,
    `step(actions[B, nj] f32) -> (obs, reward f64[B], terminated bool[B], truncated bool[B], info)`, truncation at
    `max_episode_steps` followed by AUTO-RESET (the returned observation is the reset one; the last real observation and
    info go to info["final_observation"][i] / info["final_info"][i], masks in info["_final_info"]), `call(name)`,
    `unwrapped.metadata`, `close()`.  Environments are sharded across ranks by the caller (one instance per GPU).
    """

    metadata = GuidedVisionEnv.metadata

    def __init__(self, task: str, num_envs: int, num_arms: int = 3, cameras=(), max_episode_steps: int = 300,
                 device: int = 0, solver: str = "newton", solver_iterations: int = 8, warmstart: int = 2, seed: int = 0,
                 reference_rng: bool = False, observation_height: int = 480, observation_width: int = 640,
                 free_pos=None, episode_phase=None):
        """free_pos [B, nfree, 3] (optional): fixed object placements used by every reset of environment e instead of a draw --
        scripted rollouts whose actions were planned for known placements (bench.py).  episode_phase [B] (optional): steps
        environment e has already spent in its episode after the first reset(), so that episodes end staggered over the batch."""
        self.task = TASK_OF.get(task, task)
        self.cameras = list(cameras)
        _check_cameras(self.cameras)
        self.num_envs, self.num_arms = int(num_envs), num_arms
        self.num_joints = 14 if num_arms == 2 else 21
        self._max_episode_steps = int(max_episode_steps)
        self.reference_rng = reference_rng
        self._model = _model(self.task, num_arms, device)
        self._batch = capi.Batch(self._model, self.num_envs, seed=seed)
        _configure_solver(self._batch, solver, solver_iterations, warmstart)
        self.max_reward = self._model.max_reward
        self._free_joints = model_io.load_names(self.task, num_arms)["free_joint"]
        self._elapsed = np.zeros(self.num_envs, np.int64)
        self._agent = np.zeros((self.num_envs, self.num_joints), np.float32)
        self._reward = np.zeros((self.num_envs,), np.int32)
        self._status = np.zeros((self.num_envs,), np.int32)
        self._free_pos = None if free_pos is None else np.ascontiguousarray(free_pos, np.float64).reshape(self.num_envs, len(self._free_joints), 3)
        self._phase = None if episode_phase is None else np.asarray(episode_phase, np.int64).reshape(self.num_envs)
        self.blowup_resets = 0          # environments auto-reset because the step flagged a numerical blow-up (STATUS bit 0)
        self.observation_height, self.observation_width = observation_height, observation_width
        self._cam_ids = _camera_ids(self.task, num_arms, self.cameras)
        self._render_cam = _camera_ids(self.task, num_arms, [RENDER_CAMERA])
        self.single_observation_space = spaces.Dict({
            "pixels": spaces.Dict({c: spaces.Box(low=0, high=255, shape=(observation_height, observation_width, 3),
                                                 dtype=np.uint8) for c in self.cameras}),
            "agent_pos": spaces.Box(low=-np.inf, high=np.inf, shape=(self.num_joints,), dtype=np.float64)})
        self.single_action_space = spaces.Box(low=-np.inf, high=np.inf, shape=(self.num_joints,), dtype=np.float32)
        self.observation_space, self.action_space = self.single_observation_space, self.single_action_space

    @property
    def unwrapped(self):
        return self

    def _pixels(self):
        if not self.cameras:
            return {}
        img = self._batch.render(self._cam_ids, self.observation_height, self.observation_width).cpu().numpy()
        return {c: img[:, k] for k, c in enumerate(self.cameras)}

    def _obs(self, pixels=None):
        return {"pixels": self._pixels() if pixels is None else pixels, "agent_pos": self._agent.astype(np.float64)}

    def _reset_rows(self, mask):
        fp = self._free_pos
        if fp is None and self.reference_rng:   # host draws in the reference's np.random order, env by env (SyncVectorEnv loops envs)
            fp = np.zeros((self.num_envs, len(self._free_joints), 3))
            for e in np.nonzero(mask)[0]:
                fp[e] = reference_reset_draws(self.task, self._free_joints)
        self._batch.reset(mask=None if mask.all() else mask.astype(np.uint8), free_pos=fp)
        self._elapsed[mask] = 0

    def reset(self, seed=None, options=None):
        self._reset_rows(np.ones(self.num_envs, bool))
        if self._phase is not None:
            self._elapsed[:] = self._phase
        self._agent[:] = self._batch.get(capi.AGENT_POS).cpu().numpy()
        return self._obs(), {}

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.num_envs, self.num_joints)
        self._batch.step_host(a, SIM_PHYSICS_ENV_STEP_RATIO, self._agent, self._reward, self._status)
        self._elapsed += 1
        reward = self._reward.astype(np.float64)
        terminated = np.zeros(self.num_envs, bool)
        truncated = self._elapsed >= self._max_episode_steps
        # an environment whose state blew up numerically (STATUS bit 0; SURVEY.md 8b "Errors") ends its episode here and is
        # reset like a truncated one -- MuJoCo's analogue is the mj_checkAcc warning followed by mj_resetData
        blown = (self._status & 1) != 0
        done = truncated | blown
        if done.any():
            self.blowup_resets += int(blown.sum())
            final_obs = np.empty(self.num_envs, object)
            final_info = np.empty(self.num_envs, object)
            agent64 = self._agent.astype(np.float64)
            last_px = self._pixels()                      # the last real frames, rendered before the auto-reset
            for e in np.nonzero(done)[0]:
                final_obs[e] = {"pixels": {c: v[e].copy() for c, v in last_px.items()}, "agent_pos": agent64[e].copy()}
                final_info[e] = {"is_success": bool(self._reward[e] == self.max_reward) and not blown[e],
                                 "TimeLimit.truncated": bool(truncated[e])}
                if blown[e]:
                    final_info[e]["numerical_blowup"] = True
            info = {"final_observation": final_obs, "_final_observation": done.copy(),
                    "final_info": final_info, "_final_info": done.copy()}
            self._reset_rows(done)
            fresh = self._batch.get(capi.AGENT_POS).cpu().numpy()
            self._agent[done] = fresh[done]
            truncated = done
            if self.cameras and not done.all():           # only the reset environments show a new frame: one more render, spliced
                new_px = self._pixels()
                px = {c: np.where(done[:, None, None, None], new_px[c], last_px[c]) for c in self.cameras}
            else:
                px = None if self.cameras else {}
            return self._obs(px), reward, terminated, truncated, info
        info = {"is_success": self._reward == self.max_reward, "_is_success": np.ones(self.num_envs, bool)}
        return self._obs(), reward, terminated, truncated, info

    def call(self, name, *args, **kwargs):
        v = getattr(self, name)
        if callable(v):
            v = v(*args, **kwargs)
            return v if isinstance(v, tuple) and len(v) == self.num_envs else tuple([v] * self.num_envs)
        return tuple([v] * self.num_envs)

    def render(self):
        """tuple of 225 x 300 overhead frames, one per environment (what `env.call("render")` returns, eval.py:259-263)"""
        img = self._batch.render(self._render_cam, 225, 300).cpu().numpy()
        return tuple(img[e, 0] for e in range(self.num_envs))

    # -- teleop pose-action step (reference data_collection_scripts/sim_env.py:277-312): 23-dim actions
    # [L pos3 quat4(wxyz) grip | R pos3 quat4 grip | M pos3 quat4]; GradIK for the two manipulators (sim_env.py:89-124),
    # DiffIK for the camera arm (125-138), gripper ctrl = unnorm(1 - g); everything stays on the device
    def step_pose(self, actions):
        import torch
        from . import kinematics

        if self.num_arms != 3:
            raise ValueError("pose actions drive all three arms")
        if not hasattr(self, "_ik"):
            self._ik = (kinematics.GradIK(self._model, "left", **kinematics.GRADIK_SIM),
                        kinematics.GradIK(self._model, "right", **kinematics.GRADIK_SIM),
                        kinematics.DiffIK(self._model, "middle", **kinematics.DIFFIK_SIM))
        dev = self._batch.dev
        a = torch.as_tensor(np.asarray(actions, np.float32) if not torch.is_tensor(actions) else actions,
                            dtype=torch.float32, device=dev).reshape(self.num_envs, 23)
        q = self._batch.get(capi.AGENT_POS)                      # joint positions (grippers normalised; not used by the IK)
        left = self._ik[0].run(q[:, 0:6].contiguous(), a[:, 0:3].contiguous(), a[:, 3:7].contiguous())
        right = self._ik[1].run(q[:, 7:13].contiguous(), a[:, 8:11].contiguous(), a[:, 11:15].contiguous())
        middle = self._ik[2].run(q[:, 14:21].contiguous(), a[:, 16:19].contiguous(), a[:, 19:23].contiguous())
        joint = torch.cat([left, 1.0 - a[:, 7:8], right, 1.0 - a[:, 15:16], middle], dim=1).contiguous()
        self._batch.step(joint, SIM_PHYSICS_ENV_STEP_RATIO)
        self._agent[:] = self._batch.get(capi.AGENT_POS).cpu().numpy()
        self._reward[:] = self._batch.get(capi.REWARD).cpu().numpy()
        self._elapsed += 1
        # the teleop env reports reward 0 / terminated False / truncated False on every step (sim_env.py:306-311)
        return self.get_teleop_obs(), np.zeros(self.num_envs), np.zeros(self.num_envs, bool), np.zeros(self.num_envs, bool), {}

    # the teleop environment's observation (reference data_collection_scripts/sim_env.py:160-218), batched over the leading axis
    TELEOP_CAMERAS = {"zed_cam": ("zed_cam_left", "zed_cam_right"), "cam_left_wrist": ("wrist_cam_left",),
                      "cam_right_wrist": ("wrist_cam_right",), "cam_high": ("overhead_cam",), "cam_low": ("worms_eye_cam",)}

    def get_teleop_obs(self, cameras=()):
        """{'joints': {'position' [B,21], 'velocity' [B,21]}, 'qpos' [B,nq], 'control' [B,21], 'poses': {'left','right','middle'
        [B,7] = position + wxyz quaternion of FK(ctrl)}, 'images': {...}} -- gripper position / ctrl normalised to [0, 1], gripper
        velocity divided by the ctrl range, the commanded (ctrl) end-effector poses rather than the measured ones, 'zed_cam' = the
        two 720 x 720 eye views side by side, the other cameras 480 x 640 (sim_env.py:160-218)."""
        from . import kinematics, transform_utils

        if self.num_arms != 3:
            raise ValueError("the teleop observation covers all three arms")
        m = self._model
        qpos = self._batch.get(capi.QPOS).cpu().numpy().astype(np.float64)
        qvel = self._batch.get(capi.QVEL).cpu().numpy().astype(np.float64)
        ctrl = self._batch.get(capi.CTRL).cpu().numpy().astype(np.float64)
        qadr, dadr = m.table("obs_qadr"), m.table("act_dof")
        lo, hi = m.table("act_ctrl_lo").astype(np.float64), m.table("act_ctrl_hi").astype(np.float64)
        pos, vel, con = qpos[:, qadr[:21]], qvel[:, dadr[:21]], ctrl[:, :21].copy()
        for g in (6, 13):
            pos[:, g] = (pos[:, g] - lo[g]) / (hi[g] - lo[g])
            vel[:, g] = vel[:, g] / (hi[g] - lo[g])
            con[:, g] = (con[:, g] - lo[g]) / (hi[g] - lo[g])
        poses = {}
        for arm, sl in (("left", slice(0, 6)), ("right", slice(7, 13)), ("middle", slice(14, 21))):
            if not hasattr(self, "_fk"):
                self._fk = {}
            fk = self._fk.setdefault(arm, kinematics.create_fk_fn(m, arm))
            T = np.asarray(fk(ctrl[:, sl])).reshape(-1, 4, 4)
            quat = np.asarray(transform_utils.mat2quat(np.ascontiguousarray(T[:, :3, :3]))).reshape(-1, 4)
            poses[arm] = np.concatenate([T[:, :3, 3], transform_utils.xyzw_to_wxyz(quat)], axis=1)
        images = {}
        for cam in cameras:
            key = next((k for k in self.TELEOP_CAMERAS if k in cam), None)
            if key is None:
                raise NotImplementedError(f"Camera {cam} not implemented")
            ids = _camera_ids(self.task, self.num_arms, list(self.TELEOP_CAMERAS[key]))
            h, w = (720, 720) if key == "zed_cam" else (480, 640)
            img = self._batch.render(ids, h, w).cpu().numpy()
            images[key] = np.concatenate([img[:, 0], img[:, 1]], axis=2) if key == "zed_cam" else img[:, 0]
        return {"joints": {"position": pos, "velocity": vel}, "qpos": qpos, "control": con, "poses": poses, "images": images}

    # -- device-resident variants: nothing below copies to the host.  The unchanged lerobot loop pays, per
    # step, a D2H of every frame plus a CPU uint8 -> fp32 cast (eval.py:150, utils.py:37-50); a policy that lives on the same
    # GPU can read the frames where avsim_render wrote them (av_aloha_b200/observation.py turns them into its input planes).
    def observation_device(self, render=True):
        """{"pixels": uint8 CUDA [B, ncam, H, W, 3] (None when no cameras or render=False), "agent_pos": f32 CUDA [B, nj]};
        camera k of the pixel tensor is self.cameras[k]."""
        px = None
        if self.cameras and render:
            self._px_dev = self._batch.render(self._cam_ids, self.observation_height, self.observation_width,
                                              out=getattr(self, "_px_dev", None))
            px = self._px_dev
        return {"pixels": px, "agent_pos": self._batch.get(capi.AGENT_POS)}

    def reset_device(self, render=True):
        self._reset_rows(np.ones(self.num_envs, bool))
        return self.observation_device(render), {}

    def step_device(self, actions, render=True):
        """`step` with CUDA tensors in and out: actions f32 [B, nj] on the batch's device.  Returns (observation_device(),
        reward i32 CUDA [B], terminated bool CUDA [B], truncated bool CUDA [B], info); on a step where episodes end (TimeLimit,
        then auto-reset as in `step`) info["final_success"] holds the per-env is_success of the finished episodes as a bool
        CUDA tensor (False elsewhere) and info["_final_info"] the mask -- the content of gymnasium's info["final_info"]."""
        import torch

        dev = self._batch.dev
        a = torch.as_tensor(actions, dtype=torch.float32, device=dev).reshape(self.num_envs, self.num_joints).contiguous()
        self._batch.step(a, SIM_PHYSICS_ENV_STEP_RATIO)
        self._elapsed += 1
        reward = self._batch.get(capi.REWARD)
        truncated = self._elapsed >= self._max_episode_steps           # host-side step counters: no device sync
        info = {}
        if truncated.any():
            tr = torch.as_tensor(truncated, device=dev)
            info = {"final_success": self._batch.get(capi.SUCCESS).bool() & tr, "_final_info": tr}
            self._reset_rows(truncated)
        else:
            tr = torch.zeros(self.num_envs, dtype=torch.bool, device=dev)
        return self.observation_device(render), reward, torch.zeros_like(tr), tr, info

    def success_and_max_reward(self):
        """Per-env (reward == max_reward, reward) of the last step as CUDA tensors: the payload of the one collective of the
        path (rank-sharded rollouts all_gather these at episode end)."""
        return self._batch.get(capi.SUCCESS), self._batch.get(capi.REWARD)

    def close(self):
        if getattr(self, "_batch", None) is not None:
            self._batch.close()
            self._batch = None
