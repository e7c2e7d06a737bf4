"""Generate tests/golden/ik_golden.npz by RUNNING THE REFERENCE'S OWN IK CODE (unmodified) in this container.

The reference files data_collection_scripts/{transform_utils,kinematics,diff_ik,grad_ik}.py are imported from
/root/reference as they are.  They need two things that are absent offline:
  * `import mujoco`          -> a stub module whose mj_kinematics(model, data) is a no-op;
  * a dm_control `physics`   -> a fake whose .bind(joints) yields {qpos, xaxis, xanchor, range} and .bind(site) yields
                                {xmat, xpos}, filled with the q = 0 screw axes / anchors / site pose that OUR model
                                compiler derived from aloha_sim.xml (arrays ik_w0, ik_p0, ik_site0, ik_range of the
                                compiled .avm).  So the golden vectors pin the IK *algorithms* (PoE FK, space Jacobian,
                                damped least squares + null space, finite-difference descent); the q = 0 geometry is
                                pinned separately against the XML itself: tests/test_host_logic.py compares it with an independent walk
                                of aloha_sim.xml (tools/gen_ik_geometry_golden.py -> tests/golden/ik_geometry.json).
The reference cannot travel to the GPU box, so the outputs are committed as a small fixture.

    python tools/gen_ik_golden.py           # ~2 min (numba cold JIT of the reference closures)
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/data_collection_scripts"
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")

mj = types.ModuleType("mujoco")
mj.mj_kinematics = lambda m, d: None
sys.modules["mujoco"] = mj
sys.path.insert(0, REF)

from av_aloha_b200 import model_io  # noqa: E402

import diff_ik as ref_diff_ik  # noqa: E402
import grad_ik as ref_grad_ik  # noqa: E402
import kinematics as ref_kin  # noqa: E402
import transform_utils as ref_tu  # noqa: E402


class _Obj:
    pass


class FakeBind:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class FakePhysics:
    """Duck-typed stand-in for dm_control's Physics, holding the q = 0 kinematic data of one arm."""

    def __init__(self, w0, p0, site0, rng):
        self.model = _Obj(); self.model.ptr = None
        self.data = _Obj(); self.data.ptr = None
        self._j = FakeBind(qpos=np.zeros(len(w0)), xaxis=w0.copy(), xanchor=p0.copy(), range=rng.copy())
        self._s = FakeBind(xmat=site0[:3, :3].reshape(9).copy(), xpos=site0[:3, 3].copy())

    def bind(self, what):
        return self._j if isinstance(what, list) else self._s


def arm_physics(avm, arm):
    n = int(avm["ik_ndof"][arm])
    return FakePhysics(avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm], avm["ik_range"][arm, :n]), n


def f32(x):
    """Inputs are rounded to float32-representable values BEFORE the reference sees them: the C-ABI takes fp32 I/O, so the
    reference and the kernels get bit-identical inputs (the reference still computes in float64)."""
    return np.asarray(x, np.float32).astype(np.float64)


def random_quat_wxyz(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def main():
    avm = model_io.load_avm(model_io.model_path("slot_insertion", 3))
    rng = np.random.default_rng(20240917)
    out = {}
    HOME = {0: np.array([0, -0.082, 1.06, 0, -0.953, 0]), 1: np.array([0, -0.082, 1.06, 0, -0.953, 0]),
            2: np.array([0, -0.8, 0.8, 0, 0.5, 0, 0])}
    for arm in range(3):
        phys, n = arm_physics(avm, arm)
        joints = list(range(n))
        fk = ref_kin.create_fk_fn(phys, joints, "site")
        jac = ref_kin.create_jac_fn(phys, joints)
        lo, hi = avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1]
        # ---- FK / Jacobian: the reference's own self-check pose q = [1,0,0,-1,0,1] (kinematics.py:148) + random q in range
        N = 64
        q = rng.uniform(lo, hi, size=(N, n))
        q[0, :6] = [1, 0, 0, -1, 0, 1]
        q[1] = 0.0
        q[2] = HOME[arm]
        q = f32(q)
        T = np.stack([fk(q[i].copy()) for i in range(N)])
        J = np.stack([jac(q[i].copy()) for i in range(N)])
        out[f"fk_q_{arm}"], out[f"fk_T_{arm}"], out[f"jac_J_{arm}"] = q, T, J
        # ---- DiffIK (sim parameters sim_env.py:125-138; the real-robot set real_env.py:84-97 as a second case)
        for tag, kw in (("sim", dict(k_pos=0.9, k_ori=0.9, integration_dt=0.04)),
                        ("real", dict(k_pos=0.3, k_ori=0.3, integration_dt=0.02))):
            k_null = np.array([10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0])[:n]
            ctl = ref_diff_ik.DiffIK(physics=phys, joints=joints, actuators=joints, eef_site="site", damping=1.0e-4,
                                     k_null=k_null, q0=HOME[arm].astype(np.float64), max_angvel=3.14, iterations=10, **kw)
            ctl1 = ref_diff_ik.DiffIK(physics=phys, joints=joints, actuators=joints, eef_site="site", damping=1.0e-4,
                                      k_null=k_null, q0=HOME[arm].astype(np.float64), max_angvel=3.14, iterations=1, **kw)
            Nd = 48
            qs = f32(np.clip(HOME[arm] + rng.normal(0, 0.35, size=(Nd, n)), lo, hi))
            tgt_q = np.clip(qs + rng.normal(0, 0.15, size=(Nd, n)), lo, hi)
            pos = np.zeros((Nd, 3)); quat = np.zeros((Nd, 4)); res = np.zeros((Nd, n)); res1 = np.zeros((Nd, n))
            for i in range(Nd):
                Tt = fk(tgt_q[i].copy())
                pos[i] = Tt[:3, 3]
                quat[i] = ref_tu.xyzw_to_wxyz(ref_tu.mat2quat(Tt[:3, :3].copy()))
                if i % 4 == 3:   # some targets that are not exactly reachable poses
                    pos[i] += rng.normal(0, 0.05, 3)
                    quat[i] = random_quat_wxyz(rng)
                pos[i], quat[i] = f32(pos[i]), f32(quat[i])
                res[i] = ctl.run(qs[i].copy(), pos[i].copy(), quat[i].copy())
                res1[i] = ctl1.run(qs[i].copy(), pos[i].copy(), quat[i].copy())      # the single-iteration map
            out[f"diffik_{tag}_out1_{arm}"] = res1
            out[f"diffik_{tag}_q_{arm}"], out[f"diffik_{tag}_pos_{arm}"] = qs, pos
            out[f"diffik_{tag}_quat_{arm}"], out[f"diffik_{tag}_out_{arm}"] = quat, res
        # ---- GradIK (sim parameters sim_env.py:89-124); 6-dof weights padded for the 7-dof arm
        cw = np.array([10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:n]
        ctl = ref_grad_ik.GradIK(physics=phys, joints=joints, actuators=joints, eef_site="site", step_size=0.0001,
                                 min_cost_delta=1.0e-12, max_iterations=50, position_weight=500.0, rotation_weight=100.0,
                                 joint_center_weight=cw, joint_displacement_weight=np.array(n * [50.0]),
                                 position_threshold=0.001, rotation_threshold=0.001, max_pos_diff=0.1, max_rot_diff=0.3,
                                 joint_p=0.9)
        import copy
        ctl8 = copy.copy(ctl)
        ctl8.max_iterations = 8          # the first iterations pin the algorithm tightly (see tests/test_ik_golden.py)
        Ng = 32
        qs = f32(np.clip(HOME[arm] + rng.normal(0, 0.3, size=(Ng, n)), lo, hi))
        tgt_q = np.clip(qs + rng.normal(0, 0.12, size=(Ng, n)), lo, hi)
        pos = np.zeros((Ng, 3)); quat = np.zeros((Ng, 4)); res = np.zeros((Ng, n)); res8 = np.zeros((Ng, n))
        lim_pos = np.zeros((Ng, 3)); lim_mat = np.zeros((Ng, 3, 3)); best_cost = np.zeros(Ng)
        for i in range(Ng):
            Tt = fk(tgt_q[i].copy())
            pos[i] = Tt[:3, 3]
            quat[i] = ref_tu.xyzw_to_wxyz(ref_tu.mat2quat(Tt[:3, :3].copy()))
            if i % 4 == 3:       # far targets: exercise limit_pose
                pos[i] += rng.normal(0, 0.2, 3)
                quat[i] = random_quat_wxyz(rng)
            pos[i], quat[i] = f32(pos[i]), f32(quat[i])
            res[i] = ctl.run(qs[i].copy(), pos[i].copy(), quat[i].copy())
            res8[i] = ctl8.run(qs[i].copy(), pos[i].copy(), quat[i].copy())
            # the limited target and the cost of the reference's best iterate (best = q_start + (out - q_start) / joint_p)
            Tc = fk(qs[i].copy())
            lp, lm = ref_tu.limit_pose(Tc[:3, 3].copy(), Tc[:3, :3].copy(), pos[i].copy(),
                                       ref_tu.quat2mat(ref_tu.wxyz_to_xyzw(quat[i].copy())), 0.1, 0.3)
            lim_pos[i], lim_mat[i] = lp, lm
            best = qs[i] + (res[i] - qs[i]) / 0.9
            best_cost[i] = ctl.cost_fn(best, qs[i].copy(), lp, np.ascontiguousarray(lm))
        out[f"gradik_out8_{arm}"], out[f"gradik_limpos_{arm}"], out[f"gradik_limmat_{arm}"] = res8, lim_pos, lim_mat
        out[f"gradik_bestcost_{arm}"] = best_cost
        out[f"gradik_q_{arm}"], out[f"gradik_pos_{arm}"], out[f"gradik_quat_{arm}"], out[f"gradik_out_{arm}"] = qs, pos, quat, res
        print(f"arm {arm}: done", flush=True)
    # ---- transform_utils primitives (a13)
    Nq = 32
    quats = np.stack([random_quat_wxyz(rng) for _ in range(Nq)])            # treated as xyzw by quat2mat
    mats = np.stack([ref_tu.quat2mat(quats[i].copy()) for i in range(Nq)])
    back = np.stack([ref_tu.mat2quat(mats[i].copy()) for i in range(Nq)])
    aa = np.stack([ref_tu.quat2axisangle(quats[i].copy()) for i in range(Nq)])
    aq = np.stack([ref_tu.axisangle2quat(aa[i].copy()) for i in range(Nq)])
    ang = np.stack([ref_tu.angular_error(mats[i].copy(), mats[(i + 1) % Nq].copy()) for i in range(Nq)])
    out.update(tu_quat_xyzw=quats, tu_mat=mats, tu_mat2quat=back, tu_axisangle=aa, tu_axisangle2quat=aq, tu_angerr=ang)
    dst = os.path.join(ROOT, "tests", "golden", "ik_golden.npz")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
