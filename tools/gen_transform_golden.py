"""tests/golden/transform_golden.npz: the reference's OWN transform_utils.py functions (data_collection_scripts/transform_utils.py,
imported unmodified, numba-compiled here) evaluated on seeded random inputs, plus DiffIK at kinematic singularities (reference
diff_ik.py run unmodified behind the mujoco stub of tools/gen_ik_golden.py) -- the np.linalg.pinv branch of its null-space term.

    python tools/gen_transform_golden.py        # ~1 min (numba JIT)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_ik_golden as g   # sets up the mujoco stub, sys.path and imports the reference modules  # noqa: E402

tu = g.ref_tu
rng = np.random.default_rng(77)
N = 96
out = {}


def rand_rot():
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    return tu.quat2mat(q.copy()), q


R = np.stack([rand_rot()[0] for _ in range(N)])
Q = np.stack([rand_rot()[1] for _ in range(N)])
Q[0] = [0, 0, 0, 1]; Q[1] = [0, 0, 1e-9, 1]; Q[2] = [1, 0, 0, 0]
out["mat"] = R
out["mat2quat"] = np.stack([tu.mat2quat(np.ascontiguousarray(R[i])) for i in range(N)])
out["quat"] = Q
out["quat2mat"] = np.stack([tu.quat2mat(Q[i].copy()) for i in range(N)])
out["quat2axisangle"] = np.stack([tu.quat2axisangle(Q[i].copy()) for i in range(N)])
V = rng.normal(0, 1.2, size=(N, 3)); V[0] = 0; V[1] = [1e-10, 0, 0]
out["axisangle"] = V
out["axisangle2quat"] = np.stack([tu.axisangle2quat(V[i].copy()) for i in range(N)])
R2 = np.stack([rand_rot()[0] for _ in range(N)])
out["mat_b"] = R2
out["angular_error"] = np.stack([tu.angular_error(np.ascontiguousarray(R[i]), np.ascontiguousarray(R2[i])) for i in range(N)])
cp, tp = rng.normal(0, 0.3, size=(N, 3)), None
tp = cp + rng.normal(0, 0.12, size=(N, 3))
# targets near the current orientation (so that some stay inside max_rot_diff) and far from it
Rt = np.stack([tu.quat2mat(tu.axisangle2quat(rng.normal(0, 0.25 if i % 2 else 1.0, 3))) @ R[i] for i in range(N)])
out["cur_pos"], out["tgt_pos"], out["tgt_mat"] = cp, tp, Rt
lp = [tu.limit_pose(cp[i].copy(), np.ascontiguousarray(R[i]), tp[i].copy(), np.ascontiguousarray(Rt[i]), 0.1, 0.3) for i in range(N)]
out["limit_pos"], out["limit_mat"] = np.stack([p for p, m in lp]), np.stack([m for p, m in lp])
out["within"] = np.array([tu.within_pose_threshold(cp[i].copy(), np.ascontiguousarray(R[i]), tp[i].copy(), np.ascontiguousarray(Rt[i]), 0.15, 0.4)
                          for i in range(N)])
W = rng.normal(size=(N, 3)); W /= np.linalg.norm(W, axis=1, keepdims=True)
Vv = rng.normal(0, 0.5, size=(N, 3)); TH = rng.uniform(-3, 3, N)
W[0] = 0; Vv[0] = [0, 0, 1.0]          # prismatic screw
out["exp_w"], out["exp_v"], out["exp_theta"] = W, Vv, TH
out["exp2mat"] = np.stack([tu.exp2mat(W[i].copy(), Vv[i].copy(), TH[i]) for i in range(N)])
T = out["exp2mat"]
out["adjoint"] = np.stack([tu.adjoint(np.ascontiguousarray(T[i])) for i in range(N)])
out["mat2pose_quat"] = np.stack([tu.mat2pose(np.ascontiguousarray(T[i]))[1] for i in range(N)])

# ---- DiffIK at singular configurations: wrist_angle = 0 puts forearm_roll and wrist_rotate on one line (rank 5); one iteration
avm = g.model_io.load_avm(g.model_io.model_path("slot_insertion", 3))
HOME = {0: np.array([0, -0.082, 1.06, 0, -0.953, 0]), 2: np.array([0, -0.8, 0.8, 0, 0.5, 0, 0])}
for arm in (0, 2):
    phys, n = g.arm_physics(avm, arm)
    joints = list(range(n))
    fk = g.ref_kin.create_fk_fn(phys, joints, "site")
    k_null = np.array([10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0])[:n]
    ctl = g.ref_diff_ik.DiffIK(physics=phys, joints=joints, actuators=joints, eef_site="site", damping=1.0e-4, k_null=k_null,
                               q0=HOME[arm].astype(np.float64), max_angvel=3.14, iterations=1, k_pos=0.9, k_ori=0.9, integration_dt=0.04)
    Ns = 12
    qs = np.clip(HOME[arm] + rng.normal(0, 0.3, size=(Ns, n)), avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1])
    qs[:, 4] = 0.0                       # the singular wrist
    if arm == 2:
        qs[::2, 5] = 0.0
    qs = g.f32(qs)
    pos, quat, res = np.zeros((Ns, 3)), np.zeros((Ns, 4)), np.zeros((Ns, n))
    for i in range(Ns):
        Tt = fk(np.clip(qs[i] + rng.normal(0, 0.1, n), avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1]))
        pos[i] = g.f32(Tt[:3, 3]); quat[i] = g.f32(tu.xyzw_to_wxyz(tu.mat2quat(Tt[:3, :3].copy())))
        res[i] = ctl.run(qs[i].copy(), pos[i].copy(), quat[i].copy())
    out[f"sing_q_{arm}"], out[f"sing_pos_{arm}"], out[f"sing_quat_{arm}"], out[f"sing_out_{arm}"] = qs, pos, quat, res
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "transform_golden.npz"), **out)
print("wrote tests/golden/transform_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("sing_out")})
