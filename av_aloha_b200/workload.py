"""Synthetic scripted-policy workload for SlotInsertion (bench.py, tests): joint-target action streams that actually
perform the task -- left hand reaches and pinches a slot rail, right hand grasps the stick, lifts it, carries it over
the slot and lowers it into the gap -- for per-environment object placements drawn from the reference's reset ranges
(gym_guided_vision/env.py:517-533).  This is SURVEY.md 8(d) config 2's "IK-solved offline for the sampled stick pose".

Pure numpy, no GPU, no reference import: the waypoint IK is a small batched damped-least-squares solver on the
product-of-exponentials model of the arms (screw axes from the compiled model), so the same stream feeds the CUDA arm
and the CPU baseline arm.  It is workload generation, not part of the product path.
"""
from __future__ import annotations

import numpy as np

from . import model_io

HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float64)
PAD_FWD, PAD_DOWN = 0.145, 0.006     # finger-pad centre in the gripper frame: 14.5 cm ahead, 6 mm below the EE site


def _rodrigues(w, th):
    """R = exp([w] th) for unit axes w[..., 3], angles th[...]"""
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    s, c = np.sin(th)[..., None, None], np.cos(th)[..., None, None]
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def fk_jac(q, w0, p0, site0):
    """Batched PoE forward kinematics + geometric Jacobian of the EE site.  q [B, n] -> R [B,3,3], p [B,3], J [B,6,n]."""
    B, n = q.shape
    R = np.tile(np.eye(3), (B, 1, 1))
    t = np.zeros((B, 3))
    axes, anchors = [], []
    for i in range(n):
        a = R @ w0[i]                                   # joint axis and anchor in the world at the current configuration
        r = (R @ p0[i]) + t
        axes.append(a); anchors.append(r)
        Ri = _rodrigues(np.broadcast_to(w0[i], (B, 3)), q[:, i])
        # T <- T * exp(S_i q_i): rotation about the line (w0_i, p0_i)
        t = t + np.einsum("bij,bj->bi", R, p0[i] - np.einsum("bij,j->bi", Ri, p0[i]))
        R = R @ Ri
    Re = R @ site0[:3, :3]
    pe = np.einsum("bij,j->bi", R, site0[:3, 3]) + t
    J = np.zeros((B, 6, n))
    for i in range(n):
        J[:, :3, i] = np.cross(axes[i], pe - anchors[i])
        J[:, 3:, i] = axes[i]
    return Re, pe, J


def solve_ik(q0, p_t, R_t, w0, p0, site0, lo, hi, iters=60, damping=1e-3):
    """Damped least squares to the pose (p_t [B,3], R_t [B,3,3]) from q0 [B,n]; returns q [B,n] and final errors."""
    q = q0.copy()
    for _ in range(iters):
        R, p, J = fk_jac(q, w0, p0, site0)
        ew = 0.5 * sum(np.cross(R[:, :, k], R_t[:, :, k]) for k in range(3))
        err = np.concatenate([p_t - p, ew], axis=1)
        A = J @ J.transpose(0, 2, 1) + damping * np.eye(6)
        dq = np.einsum("bij,bi->bj", J, np.linalg.solve(A, err[..., None])[..., 0])
        q = np.clip(q + np.clip(dq, -0.3, 0.3), lo, hi)
    R, p, _ = fk_jac(q, w0, p0, site0)
    return q, np.linalg.norm(p_t - p, axis=1)


def _roty(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def sample_object_positions(B, seed):
    """[B, 2, 3] (slot, stick) positions from the reference's uniform ranges (env.py:517-533)."""
    rng = np.random.default_rng(seed)
    slot = np.stack([rng.uniform(-0.05, 0.05, B), rng.uniform(0.1, 0.15, B), np.zeros(B)], 1)
    stick = np.stack([rng.uniform(-0.08, 0.08, B), rng.uniform(-0.1, 0.0, B), np.zeros(B)], 1)
    return np.stack([slot, stick], 1)


def slot_insertion_script(T, obj_pos, seed, pitch=1.0, noise=0.01):
    """Action stream [T, B, 21] float32 for SlotInsertion-3Arms and the matching object positions.

    Phases (25 Hz): 0-50 move to pre-grasp, 50-80 descend, 80-105 close, 105-150 lift the stick, 150-215 carry it over
    the slot, 215-260 lower it into the gap, then hold.  The active-vision (middle) arm sweeps slowly.  Per-env N(0, noise)
    joint noise de-synchronises the environments.
    """
    B = len(obj_pos)
    avm = model_io.load_avm(model_io.model_path("slot_insertion", 3))
    rng = np.random.default_rng(seed)
    slot, stick = obj_pos[:, 0], obj_pos[:, 1]
    arms = {}
    for arm, sgn in ((0, 1.0), (1, -1.0)):                # left arm points +x, right arm points -x
        n = int(avm["ik_ndof"][arm])
        w0, p0, site0 = avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm]
        lo, hi = avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1]
        Rt = _roty(sgn * pitch) @ site0[:3, :3]
        off = _roty(sgn * pitch) @ np.array([sgn * PAD_FWD, 0.0, -PAD_DOWN])
        arms[arm] = (w0, p0, site0, lo, hi, Rt, off)

    def ik(arm, pad_xyz, q_init):
        w0, p0, site0, lo, hi, Rt, off = arms[arm]
        q, err = solve_ik(q_init, pad_xyz - off, np.broadcast_to(Rt, (B, 3, 3)), w0, p0, site0, lo, hi)
        return q

    # left hand: the near rail of the slot (slot-2 at y - 0.032), pinched 6 cm from the slot's left end
    rail = slot + np.array([-0.04, -0.032, 0.0])
    qL0 = np.tile(HOME[:6], (B, 1))
    qL_pre = ik(0, rail + [0, 0, 0.10], qL0)
    qL_grasp = ik(0, rail + [0, 0, 0.028], qL_pre)
    # right hand: the stick, 9 cm right of its centre
    grip = stick + np.array([0.09, 0.0, 0.0])
    qR0 = np.tile(HOME[7:13], (B, 1))
    qR_pre = ik(1, grip + [0, 0, 0.10], qR0)
    qR_grasp = ik(1, grip + [0, 0, 0.022], qR_pre)
    qR_lift = ik(1, grip + [0, 0, 0.13], qR_grasp)
    over = np.stack([slot[:, 0] + 0.09 + 0.02, slot[:, 1], np.full(B, 0.13)], 1)     # stick centred over the gap
    qR_over = ik(1, over, qR_lift)
    qR_in = ik(1, over - [0, 0, 0.09], qR_over)

    def lerp(a, b, s):
        s = np.clip(s, 0.0, 1.0)
        s = s * s * (3 - 2 * s)
        return a + s * (b - a)

    jn = rng.normal(0.0, noise, size=(B, 21))
    jn[:, [6, 13]] = 0.0
    sweep = rng.uniform(-0.25, 0.25, size=(B, 1))
    acts = np.empty((T, B, 21), np.float32)
    for t in range(T):
        a = np.tile(HOME, (B, 1))
        if t < 50:
            qL, qR, gL, gR = lerp(qL0, qL_pre, t / 50), lerp(qR0, qR_pre, t / 50), 1.0, 1.0
        elif t < 80:
            qL, qR, gL, gR = lerp(qL_pre, qL_grasp, (t - 50) / 30), lerp(qR_pre, qR_grasp, (t - 50) / 30), 1.0, 1.0
        elif t < 105:
            g = 1.0 - (t - 80) / 25
            qL, qR, gL, gR = qL_grasp, qR_grasp, g, g
        elif t < 150:
            qL, qR, gL, gR = qL_grasp, lerp(qR_grasp, qR_lift, (t - 105) / 45), 0.0, 0.0
        elif t < 215:
            qL, qR, gL, gR = qL_grasp, lerp(qR_lift, qR_over, (t - 150) / 65), 0.0, 0.0
        elif t < 260:
            qL, qR, gL, gR = qL_grasp, lerp(qR_over, qR_in, (t - 215) / 45), 0.0, 0.0
        else:
            qL, qR, gL, gR = qL_grasp, qR_in, 0.0, 0.0
        a[:, 0:6], a[:, 7:13] = qL, qR
        a[:, 14] = HOME[14] + sweep[:, 0] * np.sin(2 * np.pi * t / T)
        a += jn
        a[:, 6], a[:, 13] = gL, gR
        acts[t] = a
    return acts
