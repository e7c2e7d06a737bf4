"""transform_utils primitives one by one against the reference's own numba functions.

tests/golden/transform_golden.npz = outputs of data_collection_scripts/transform_utils.py (imported unmodified,
tools/gen_transform_golden.py) on seeded inputs, plus DiffIK at kinematic singularities (the np.linalg.pinv branch of
diff_ik.py:72).  CPU leg: `avsim_transform_kernel` / `avsim_diffik_kernel` source through the warp emulator; GPU leg: the same
through the C-ABI (`av_aloha_b200.transform_utils`, `kinematics.DiffIK` with the reference's constructor signature).
Tolerance 1e-12 for the closed forms (both sides fp64); mat2quat 1e-7 (the reference takes an eigenvector of a 4x4, the kernel
the closed form; the test matrices come out of the reference's quat2mat, whose float32 round trip leaves them orthogonal to 1e-8
only, and the two constructions split that error differently); quat2mat 5e-7: the reference casts the quaternion to float32 and
numba compiles the products with fastmath, so its result is a float32 computation whose last bit depends on LLVM's contraction
choices -- the kernel does the same float32 round trip, not the same instruction sequence.
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "transform_golden.npz"))
OPS = dict(mat2quat=0, quat2mat=1, quat2axisangle=2, axisangle2quat=3, angular_error=4, limit_pose=5, exp2mat=6, adjoint=7, within=8)


def _pose(p, m):
    return np.concatenate([p.reshape(-1, 3), m.reshape(-1, 9)], axis=1)


def _check(run):
    """run(op, a, wa, wout, b=None, p0=0, p1=0) -> [n, wout]"""
    q = run(OPS["mat2quat"], Z["mat"], 9, 4)
    assert np.abs(q - Z["mat2quat"]).max() <= 1e-7 and (q[:, 3] >= 0).all()
    assert np.abs(run(OPS["quat2mat"], Z["quat"], 4, 9).reshape(-1, 3, 3) - Z["quat2mat"]).max() <= 5e-7     # float32 inside, see header
    assert np.abs(run(OPS["quat2axisangle"], Z["quat"], 4, 3) - Z["quat2axisangle"]).max() <= 1e-12
    assert np.abs(run(OPS["axisangle2quat"], Z["axisangle"], 3, 4) - Z["axisangle2quat"]).max() <= 1e-12
    assert np.abs(run(OPS["angular_error"], Z["mat"], 9, 3, b=Z["mat_b"]) - Z["angular_error"]).max() <= 1e-12
    lp = run(OPS["limit_pose"], _pose(Z["cur_pos"], Z["mat"]), 12, 12, b=_pose(Z["tgt_pos"], Z["tgt_mat"]), p0=0.1, p1=0.3)
    assert np.abs(lp[:, :3] - Z["limit_pos"]).max() <= 1e-12
    assert np.abs(lp[:, 3:].reshape(-1, 3, 3) - Z["limit_mat"]).max() <= 1e-6      # through mat2quat -> axis-angle -> float32 quat2mat
    assert (np.linalg.norm(Z["limit_pos"] - Z["cur_pos"], axis=1) <= 0.1 + 1e-12).all()
    w = run(OPS["within"], _pose(Z["cur_pos"], Z["mat"]), 12, 1, b=_pose(Z["tgt_pos"], Z["tgt_mat"]), p0=0.15, p1=0.4)
    assert np.array_equal(w.reshape(-1) > 0.5, Z["within"]) and 0 < Z["within"].sum() < len(Z["within"])
    x = np.concatenate([Z["exp_w"], Z["exp_v"], Z["exp_theta"][:, None]], axis=1)
    T = run(OPS["exp2mat"], x, 7, 16).reshape(-1, 4, 4)
    assert np.abs(T - Z["exp2mat"]).max() <= 1e-12
    assert np.abs(run(OPS["adjoint"], Z["exp2mat"].reshape(-1, 16), 16, 36).reshape(-1, 6, 6) - Z["adjoint"]).max() <= 1e-12


def test_primitives_emulated_kernel_vs_reference():
    from tests.emu.emu import emu_transform
    _check(lambda op, a, wa, wout, b=None, p0=0.0, p1=0.0: emu_transform(op, a, wa, wout, b=b, p0=p0, p1=p1))


def test_diffik_at_singular_configurations_emulated_vs_reference(slot_model_path):
    """wrist_angle = 0: J has rank 5, J J' is singular; the reference's pinv drops the vanishing direction -- so does the kernel"""
    from av_aloha_b200.capi import DiffIKParams
    from tests.emu.emu import EmuBatch, emu_diffik
    eb = EmuBatch(slot_model_path, 1)
    for arm in (0, 2):
        n = Z[f"sing_q_{arm}"].shape[1]
        p = DiffIKParams()
        p.k_pos, p.k_ori, p.damping, p.max_angvel, p.integration_dt, p.iterations = 0.9, 0.9, 1e-4, 3.14, 0.04, 1
        home = [0, -0.082, 1.06, 0, -0.953, 0, 0] if arm == 0 else [0, -0.8, 0.8, 0, 0.5, 0, 0]
        for k in range(7):
            p.k_null[k] = [10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0][k] if k < n else 0.0
            p.q0[k] = home[k] if k < n else 0.0
        out = emu_diffik(eb, arm, Z[f"sing_q_{arm}"], Z[f"sing_pos_{arm}"], Z[f"sing_quat_{arm}"], p)
        assert np.isfinite(out).all()
        assert np.abs(out - Z[f"sing_out_{arm}"]).max() <= 2e-5, (arm, np.abs(out - Z[f"sing_out_{arm}"]).max())


@pytest.mark.gpu
def test_primitives_gpu_vs_reference():
    from av_aloha_b200 import transform_utils as tu

    def run(op, a, wa, wout, b=None, p0=0.0, p1=0.0):
        return tu._run(op, a, wa, wout, b=b, wb=0 if b is None else np.asarray(b).reshape(len(a), -1).shape[1], p0=p0, p1=p1)[0].reshape(-1, wout)
    _check(run)
    # the named mirror functions, single items and batches
    pos, quat = tu.mat2pose(Z["exp2mat"][5])
    assert np.abs(quat - Z["mat2pose_quat"][5]).max() <= 1e-7 and np.array_equal(pos, Z["exp2mat"][5][:3, 3])
    assert np.abs(tu.pose2mat(pos, quat)[:3, :3] - Z["exp2mat"][5][:3, :3]).max() <= 1e-6
    assert np.array_equal(tu.wxyz_to_xyzw(tu.xyzw_to_wxyz(quat)), quat)
    lp, lm = tu.limit_pose(Z["cur_pos"], Z["mat"], Z["tgt_pos"], Z["tgt_mat"], 0.1, 0.3)
    assert np.abs(lp - Z["limit_pos"]).max() <= 1e-12 and lm.shape == (len(lp), 3, 3)
    assert np.abs(tu.exp2rot(Z["exp_w"][3], Z["exp_theta"][3]) - Z["exp2mat"][3][:3, :3]).max() <= 1e-12


@pytest.mark.gpu
def test_reference_signature_ik_constructors_and_singular_diffik(slot_model_path):
    """DiffIK / GradIK / create_fk_fn built the way data_collection_scripts/sim_env.py:89-138 builds them: (physics, joints,
    actuators, eef_site, ...) with the arm's joint names; and the singular-wrist DiffIK vectors through that object"""
    from av_aloha_b200 import capi, kinematics, model_io
    model = capi.Model(slot_model_path, 0)
    names = model_io.load_names("slot_insertion", 3)["joint"]
    mid = [j for j in names if j.startswith("middle_")]
    left = [j for j in names if j.startswith("left_")][:6]
    ctl = kinematics.DiffIK(model, mid, mid, "middle_zed_camera_center", 0.9, 0.9, 1e-4, np.array([10.0, 10, 10, 10, 5, 5, 5]),
                            np.array([0, -0.8, 0.8, 0, 0.5, 0, 0]), 3.14, 0.04, 1)
    out = ctl.run(Z["sing_q_2"], Z["sing_pos_2"], Z["sing_quat_2"])
    assert np.isfinite(out).all() and np.abs(out - Z["sing_out_2"]).max() <= 2e-5
    g = kinematics.GradIK(physics=model, joints=left, actuators=left, eef_site="left_gripper_control")
    assert abs(g.params.max_rot_diff - 0.3) < 1e-7 and abs(g.params.joint_p - 0.1) < 1e-7       # the reference's class defaults
    fk = kinematics.create_fk_fn(model, left, "left_gripper_control")
    assert np.allclose(fk(np.zeros(6)), kinematics.create_fk_fn(model, "left")(np.zeros(6)))
    with pytest.raises(ValueError):
        kinematics.DiffIK(model, left[:5], None, None)
    with pytest.raises(ValueError):
        kinematics.create_fk_fn(model, left, "right_gripper_control")
