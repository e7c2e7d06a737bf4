"""Batched IK controllers behind the reference's DiffIK / GradIK interface.

Mirrors data_collection_scripts/diff_ik.py:8-22,89-90 (``DiffIK(...).run(q, target_pos, target_quat_wxyz) -> q``) and
grad_ik.py:103-166 (``GradIK(...).run(...)``), plus ``create_fk_fn`` (kinematics.py:7-26).  The reference builds numba
closures from a dm_control ``physics``; here the q = 0 screw axes / anchors / site pose live in the compiled model on
the GPU and ``run`` launches the hand-written kernels (csrc/avsim_ik.cuh) through the C-ABI for one problem or a
whole batch.  Quaternions are wxyz at this API like the reference.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi

ARM_INDEX = {"left": 0, "right": 1, "middle": 2}

# reference sim_env.py:89-138 (teleop sim env) and real_env.py:84-97
GRADIK_SIM = dict(step_size=0.0001, min_cost_delta=1.0e-12, max_iterations=50, position_weight=500.0,
                  rotation_weight=100.0, joint_center_weight=(10.0, 10.0, 1.0, 50.0, 1.0, 1.0),
                  joint_displacement_weight=(50.0,) * 6, position_threshold=0.001, rotation_threshold=0.001,
                  max_pos_diff=0.1, max_rot_diff=0.3, joint_p=0.9)
DIFFIK_SIM = dict(k_pos=0.9, k_ori=0.9, damping=1.0e-4, k_null=(10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0),
                  q0=(0, -0.8, 0.8, 0, 0.5, 0, 0), max_angvel=3.14, integration_dt=0.04, iterations=10)
DIFFIK_REAL = dict(DIFFIK_SIM, k_pos=0.3, k_ori=0.3, integration_dt=0.02)


def _arm(arm):
    return ARM_INDEX[arm] if isinstance(arm, str) else int(arm)


EEF_SITE = {"left": "left_gripper_control", "right": "right_gripper_control", "middle": "middle_zed_camera_center"}   # aloha_sim.xml:160,249,350


def _name(x):
    """name of an MJCF element the way the reference passes it (a dm_control mjcf element has .name / .full_identifier)"""
    return x if isinstance(x, str) else (getattr(x, "name", None) or getattr(x, "full_identifier", None) or str(x))


def _resolve(physics, joints, eef_site=None):
    """(capi.Model, arm index) from the reference's constructor arguments.

    The reference builds its controllers from a dm_control ``physics``, the list of MJCF joint elements of one arm and that
    arm's end-effector site (data_collection_scripts/sim_env.py:89-138: ``DiffIK(physics=..., joints=self._middle_joints,
    actuators=..., eef_site=self._middle_eef_site, ...)``).  Here ``physics`` is whatever owns the compiled model -- a
    ``capi.Model``, a ``capi.Batch``, or one of this package's environments -- and ``joints`` are the same joint elements or
    their names; the arm is read off the names ("left_waist" ... / "right_..." / "middle_...").  The kernels carry
    product-of-exponentials tables for exactly the three arm chains (first 6, 6, 7 joints), which is what every call site of the
    reference asks for; anything else raises.  For this repo's own callers ``joints`` may also be an arm name / index.
    """
    model = physics if isinstance(physics, capi.Model) else (getattr(physics, "_model", None) or getattr(physics, "model", None))
    if not isinstance(model, capi.Model):
        raise TypeError("physics must be a capi.Model, a capi.Batch or an environment of this package")
    if isinstance(joints, (str, int, np.integer)):
        return model, _arm(joints)
    names = [_name(j) for j in joints]
    if not names:
        raise ValueError("empty joint list")
    arm_name = names[0].split("_")[0]
    if arm_name not in ARM_INDEX:
        raise ValueError(f"cannot tell the arm from joint {names[0]!r}")
    arm = ARM_INDEX[arm_name]
    from . import model_io
    task = {0: "insert_peg", 1: "slot_insertion", 2: "sew_needle", 3: "tube_transfer", 4: "hook_package"}[model.task_id]
    all_j = [n for n in model_io.load_names(task, model.num_arms)["joint"] if n.startswith(arm_name + "_")]
    ndof = (6, 6, 7)[arm]
    if names != all_j[:ndof]:
        raise ValueError(f"IK tables exist for the {arm_name} arm chain {all_j[:ndof]}, got {names}")
    if eef_site is not None and _name(eef_site) != EEF_SITE[arm_name]:
        raise ValueError(f"end-effector site of the {arm_name} arm is {EEF_SITE[arm_name]!r}, got {_name(eef_site)!r}")
    return model, arm


class _Controller:
    def __init__(self, model: capi.Model, arm):
        self.model, self.arm = model, _arm(arm)
        self.lib = model.lib
        self.ndof = (6, 6, 7)[self.arm]

    def _io(self, q, pos, quat):
        import torch

        dev = torch.device("cuda", self.model.device)
        single = np.ndim(q) == 1 if not hasattr(q, "dim") else q.dim() == 1
        as_np = not hasattr(q, "is_cuda")
        t = lambda x, w: torch.as_tensor(np.asarray(x, np.float32) if not hasattr(x, "is_cuda") else x,
                                         dtype=torch.float32, device=dev).reshape(-1, w).contiguous()
        qd, pd, qtd = t(q, self.ndof), t(pos, 3), t(quat, 4)
        if not (len(qd) == len(pd) == len(qtd)):
            raise ValueError("q, target_pos and target_quat must have the same leading dimension")
        out = torch.empty_like(qd)
        return qd, pd, qtd, out, single, as_np

    @staticmethod
    def _ret(out, single, as_np):
        if as_np:
            out = out.cpu().numpy().astype(np.float64)
        return out[0] if single else out


class DiffIK(_Controller):
    """Damped-least-squares differential IK with a null-space posture term (reference diff_ik.py:51-85)."""

    def __init__(self, physics, joints="middle", actuators=None, eef_site=None, k_pos=0.9, k_ori=0.9, damping=1.0e-4,
                 k_null=DIFFIK_SIM["k_null"], q0=DIFFIK_SIM["q0"], max_angvel=3.14, integration_dt=0.04, iterations=10):
        """Reference signature (diff_ik.py:8-22): ``DiffIK(physics, joints, actuators, eef_site, k_pos, k_ori, damping, k_null,
        q0, max_angvel, integration_dt, iterations)``; see `_resolve` for what physics / joints may be."""
        model, arm = _resolve(physics, joints, eef_site)
        super().__init__(model, arm)
        self.physics, self.joints, self.actuators, self.eef_site = physics, joints, actuators, eef_site
        p = capi.DiffIKParams()
        p.k_pos, p.k_ori, p.damping, p.max_angvel, p.integration_dt = k_pos, k_ori, damping, max_angvel, integration_dt
        p.iterations = int(iterations)
        for k in range(7):
            p.k_null[k] = float(k_null[k]) if k < len(k_null) else 0.0
            p.q0[k] = float(q0[k]) if k < len(q0) else 0.0
        self.params = p

    def run(self, q, target_pos, target_quat):
        qd, pd, qtd, out, single, as_np = self._io(q, target_pos, target_quat)
        capi.check(self.lib.avsim_diffik(self.model.ptr, self.arm, C.c_void_p(qd.data_ptr()), C.c_void_p(pd.data_ptr()),
                                         C.c_void_p(qtd.data_ptr()), len(qd), C.byref(self.params),
                                         C.c_void_p(out.data_ptr()), None))
        return self._ret(out, single, as_np)


class GradIK(_Controller):
    """Finite-difference gradient-descent IK with a secant line step (reference grad_ik.py:8-99)."""

    def __init__(self, physics, joints="left", actuators=None, eef_site=None, step_size=0.0001, min_cost_delta=1.0e-12,
                 max_iterations=50, position_weight=500.0, rotation_weight=100.0,
                 joint_center_weight=GRADIK_SIM["joint_center_weight"],
                 joint_displacement_weight=GRADIK_SIM["joint_displacement_weight"], position_threshold=0.001,
                 rotation_threshold=0.001, max_pos_diff=0.1, max_rot_diff=0.3, joint_p=0.1):
        """Reference signature and defaults (grad_ik.py:101-121): ``GradIK(physics, joints, actuators, eef_site, step_size=1e-4,
        ..., max_pos_diff=0.1, max_rot_diff=0.3, joint_p=0.1)``."""
        model, arm = _resolve(physics, joints, eef_site)
        super().__init__(model, arm)
        self.physics, self.joints, self.actuators, self.eef_site = physics, joints, actuators, eef_site
        from . import model_io  # joint ranges of the arm (reference grad_ik.py:134-137 divides by the half range)

        p = capi.GradIKParams()
        p.step_size, p.min_cost_delta, p.max_iterations = step_size, min_cost_delta, int(max_iterations)
        p.position_weight, p.rotation_weight = position_weight, rotation_weight
        p.position_threshold, p.rotation_threshold = position_threshold, rotation_threshold
        p.max_pos_diff, p.max_rot_diff, p.joint_p = max_pos_diff, max_rot_diff, joint_p
        rng = model.ik_range(self.arm)
        half = 0.5 * (rng[:, 1] - rng[:, 0])
        for k in range(7):
            live = k < self.ndof and k < len(joint_center_weight)
            p.joint_center_weight[k] = float(joint_center_weight[k]) / float(half[k]) if live else 0.0
            p.joint_displacement_weight[k] = float(joint_displacement_weight[k]) if k < len(joint_displacement_weight) else 0.0
        self.params = p

    def run(self, q, target_pos, target_quat):
        qd, pd, qtd, out, single, as_np = self._io(q, target_pos, target_quat)
        capi.check(self.lib.avsim_gradik(self.model.ptr, self.arm, C.c_void_p(qd.data_ptr()), C.c_void_p(pd.data_ptr()),
                                         C.c_void_p(qtd.data_ptr()), len(qd), C.byref(self.params),
                                         C.c_void_p(out.data_ptr()), None))
        return self._ret(out, single, as_np)


def create_fk_fn(physics, joints, eef_site=None):
    """``fk(theta) -> 4x4`` (or [n,4,4]) end-effector site pose by product of exponentials.  Reference signature
    (kinematics.py:7): ``create_fk_fn(physics, joints, eef_site)``; `joints` may also be an arm name / index."""
    model, a = _resolve(physics, joints, eef_site)
    ndof = (6, 6, 7)[a]

    def forward_kinematics(theta):
        import torch

        dev = torch.device("cuda", model.device)
        as_np = not hasattr(theta, "is_cuda")
        th = torch.as_tensor(np.asarray(theta, np.float32) if as_np else theta, dtype=torch.float32, device=dev)
        single = th.dim() == 1
        th = th.reshape(-1, ndof).contiguous()
        out = torch.empty((len(th), 16), dtype=torch.float32, device=dev)
        capi.check(model.lib.avsim_fk(model.ptr, a, C.c_void_p(th.data_ptr()), len(th), C.c_void_p(out.data_ptr()), None))
        out = out.reshape(-1, 4, 4)
        if as_np:
            out = out.cpu().numpy().astype(np.float64)
        return out[0] if single else out

    return forward_kinematics


def create_jac_fn(physics, joints):
    """``jacobian(theta) -> 6 x n`` (or [N, 6, n]) space Jacobian of the end-effector site, rows [v; w].  Reference signature
    (kinematics.py:28): ``create_jac_fn(physics, joints)``; what DiffIK iterates on."""
    model, a = _resolve(physics, joints)
    ndof = (6, 6, 7)[a]

    def jacobian(theta):
        import torch

        dev = torch.device("cuda", model.device)
        as_np = not hasattr(theta, "is_cuda")
        th = torch.as_tensor(np.asarray(theta, np.float32) if as_np else theta, dtype=torch.float32, device=dev)
        single = th.dim() == 1
        th = th.reshape(-1, ndof).contiguous()
        out = torch.empty((len(th), 6, ndof), dtype=torch.float32, device=dev)
        capi.check(model.lib.avsim_jac(model.ptr, a, C.c_void_p(th.data_ptr()), len(th), C.c_void_p(out.data_ptr()), None))
        if as_np:
            out = out.cpu().numpy().astype(np.float64)
        return out[0] if single else out

    return jacobian


# ---- create_safety_fn / safety (reference data_collection_scripts/kinematics.py:54-135).  The checks are host predicates
# over the FK kernel's pose; rows are checked in the reference's order and the first failing check names the message.
SAFETY_MESSAGES = ("", "Joint tracking safety margin exceeded", "Joint limit safety margin exceeded",
                   "End effector position outside bounds", "End effector action position outside bounds",
                   "End effector pose tracking safety margin exceeded")


def _angular_error(desired, current):
    """transform_utils.py:183-194: half the sum of the column cross products, batched over the leading axis"""
    return 0.5 * sum(np.cross(current[..., :3, k], desired[..., :3, k]) for k in range(3))


def create_safety_fn(physics, joints, *args, **kwargs):
    """``safety_fn(qpos, ctrl, Taction=None) -> (ok, message)`` for one arm, or -- with a leading batch axis on qpos / ctrl
    (/ Taction) -- ``(ok bool[n], code int[n])`` with ``SAFETY_MESSAGES[code]`` the reference's message.  Reference signature
    (kinematics.py:104): ``create_safety_fn(physics, joints, eef_site, xyz_bounds, joint_limit_safety_margin=0.01,
    joint_tracking_safety_margin=1.0, eef_pos_tracking_safety_margin=0.2, eef_rot_tracking_safety_margin=3.0)``; this repo's
    own form drops eef_site: ``create_safety_fn(model, arm, xyz_bounds, ...)``."""
    names = ("xyz_bounds", "joint_limit_safety_margin", "joint_tracking_safety_margin", "eef_pos_tracking_safety_margin",
             "eef_rot_tracking_safety_margin")
    eef_site = kwargs.pop("eef_site", None)
    args = list(args)
    if args and not isinstance(joints, (str, int, np.integer)) and np.ndim(args[0]) == 0:
        eef_site = args.pop(0)                     # reference order: the site comes before the bounds
    kw = dict(joint_limit_safety_margin=0.01, joint_tracking_safety_margin=1.0, eef_pos_tracking_safety_margin=0.2,
              eef_rot_tracking_safety_margin=3.0)
    kw.update(dict(zip(names, args)))
    kw.update(kwargs)
    model, a = _resolve(physics, joints, eef_site)
    return _make_safety_fn(create_fk_fn(model, a), model.ik_range(a), kw["xyz_bounds"], kw["joint_limit_safety_margin"],
                           kw["joint_tracking_safety_margin"], kw["eef_pos_tracking_safety_margin"], kw["eef_rot_tracking_safety_margin"])


def _make_safety_fn(fk_fn, joint_range, xyz_bounds, joint_limit_safety_margin, joint_tracking_safety_margin,
                    eef_pos_tracking_safety_margin, eef_rot_tracking_safety_margin):
    """the predicates around an FK function (the CUDA kernel in the product; the CPU tests inject the emulated kernel)"""
    xyz_bounds = np.array(xyz_bounds, np.float64)
    joint_limits = np.array(joint_range, np.float64).copy()
    assert np.all(joint_limit_safety_margin < (joint_limits[:, 1] - joint_limits[:, 0]) / 2)
    joint_limits[:, 0] += joint_limit_safety_margin
    joint_limits[:, 1] -= joint_limit_safety_margin

    def safety_fn(qpos, ctrl, Taction=None):
        q, c = np.asarray(qpos, np.float64), np.asarray(ctrl, np.float64)
        single = q.ndim == 1
        q, c = np.atleast_2d(q), np.atleast_2d(c)
        T = np.asarray(fk_fn(q), np.float64).reshape(-1, 4, 4)
        code = np.zeros(len(q), np.int64)

        def flag(bad, k):                                    # the first failing check wins, as in the reference's early returns
            code[(code == 0) & bad] = k

        flag(np.any(np.abs(q - c) > joint_tracking_safety_margin, axis=1), 1)
        flag(np.any(q < joint_limits[:, 0], axis=1) | np.any(q > joint_limits[:, 1], axis=1), 2)
        flag(np.any(T[:, :3, 3] < xyz_bounds[:, 0], axis=1) | np.any(T[:, :3, 3] > xyz_bounds[:, 1], axis=1), 3)
        if Taction is not None:
            Ta = np.asarray(Taction, np.float64).reshape(-1, 4, 4)
            flag(np.any(Ta[:, :3, 3] < xyz_bounds[:, 0], axis=1) | np.any(Ta[:, :3, 3] > xyz_bounds[:, 1], axis=1), 4)
            pos_err = np.linalg.norm(Ta[:, :3, 3] - T[:, :3, 3], axis=1)
            rot_err = np.linalg.norm(_angular_error(Ta, T), axis=1)
            flag(~((pos_err < eef_pos_tracking_safety_margin) & (rot_err < eef_rot_tracking_safety_margin)), 5)
        if single:
            return bool(code[0] == 0), SAFETY_MESSAGES[int(code[0])]
        return code == 0, code

    return safety_fn
