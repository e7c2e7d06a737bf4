"""IK parity against golden vectors produced by RUNNING the reference's own numba code (tools/gen_ik_golden.py).

CPU leg  : the CUDA IK kernels' source, compiled for the host by the warp-emulation harness (tests/emu), against the
           golden vectors -- checks the kernel arithmetic in the GPU-less container.
GPU leg  : the same kernels through the C-ABI (avsim_fk / avsim_diffik / avsim_gradik) on cuda:0.
Tolerance: 1e-5 abs for FK and for ONE DiffIK iteration (the map the controller iterates; SURVEY.md 8c: the reference
           itself rounds through float32 in quat2mat, transform_utils.py:66; our I/O is fp32, arithmetic fp64).  The
           10-iteration DiffIK result runs through velocity / joint-limit clipping and can amplify last-bit differences
           (numba fastmath + SVD pinv vs our Cholesky) several-fold per iteration on far targets: median <= 1e-6,
           every case <= 2e-4.  GradIK: the first 8 descent iterations agree to 5e-6; the full 50-iteration run
           ends on a plateau where `local_cost < best_cost` is decided by the last bits of two float64 costs, so which
           plateau iterate is kept differs between any two compilations (also of the reference itself).  There the
           check is on what the controller optimises: the cost of our best iterate, evaluated independently in numpy,
           is within 3 % of the reference's best cost (the plateau iterates themselves scatter by ~1-2 % in cost), and
           the joint vectors agree to 5e-2 rad.
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ik_golden.npz")
TOL_FK, TOL_DIFFIK1, TOL_DIFFIK_MED, TOL_DIFFIK_MAX, TOL_GRADIK8, TOL_GRADIK_COST, TOL_GRADIK_Q = 1e-5, 1e-5, 1e-6, 2e-4, 5e-6, 3e-2, 5e-2


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _check_diffik(out10, out1, gold, arm, tag):
    e1 = np.abs(out1 - gold[f"diffik_{tag}_out1_{arm}"]).max(axis=1)
    e10 = np.abs(out10 - gold[f"diffik_{tag}_out_{arm}"]).max(axis=1)
    assert e1.max() <= TOL_DIFFIK1, e1.max()
    assert np.median(e10) <= TOL_DIFFIK_MED and e10.max() <= TOL_DIFFIK_MAX, (np.median(e10), e10.max())


def _diffik_params(arm, tag, iterations=10):
    from av_aloha_b200 import capi
    home = {0: [0, -0.082, 1.06, 0, -0.953, 0, 0], 1: [0, -0.082, 1.06, 0, -0.953, 0, 0], 2: [0, -0.8, 0.8, 0, 0.5, 0, 0]}[arm]
    p = capi.DiffIKParams()
    p.k_pos = p.k_ori = 0.9 if tag == "sim" else 0.3
    p.integration_dt = 0.04 if tag == "sim" else 0.02
    p.damping, p.max_angvel, p.iterations = 1.0e-4, 3.14, iterations
    for k, v in enumerate([10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0]):
        p.k_null[k] = v
        p.q0[k] = home[k]
    return p


def _gradik_cost(gold, arm, avm, best):
    """numpy restatement of cost_fn (reference grad_ik.py:176-196) on the limited target stored with the golden vectors"""
    from av_aloha_b200 import workload
    n = int(avm["ik_ndof"][arm])
    R, p, _ = workload.fk_jac(best, avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm])
    Rt, pt, q0 = gold[f"gradik_limmat_{arm}"], gold[f"gradik_limpos_{arm}"], gold[f"gradik_q_{arm}"]
    ew = 0.5 * sum(np.cross(R[:, :, k], Rt[:, :, k]) for k in range(3))
    rng = avm["ik_range"][arm, :n]
    centers, half = 0.5 * (rng[:, 0] + rng[:, 1]), 0.5 * (rng[:, 1] - rng[:, 0])
    cw = np.array([10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0])[:n] / half
    return ((500.0 * np.linalg.norm(pt - p, axis=1)) ** 2 + (100.0 * np.linalg.norm(ew, axis=1)) ** 2
            + ((cw * (best - centers)) ** 2).sum(1) + ((50.0 * (best - q0)) ** 2).sum(1))


def _check_gradik(out50, out8, gold, arm, avm):
    assert np.abs(out8 - gold[f"gradik_out8_{arm}"]).max() <= TOL_GRADIK8
    q0, ref = gold[f"gradik_q_{arm}"], gold[f"gradik_out_{arm}"]
    assert np.abs(out50 - ref).max() <= TOL_GRADIK_Q
    ref_cost = gold[f"gradik_bestcost_{arm}"]
    assert np.allclose(_gradik_cost(gold, arm, avm, q0 + (ref - q0) / 0.9), ref_cost, rtol=1e-9, atol=1e-9)   # restatement pinned
    ours = _gradik_cost(gold, arm, avm, q0 + (out50.astype(np.float64) - q0) / 0.9)
    assert (ours <= ref_cost * (1 + TOL_GRADIK_COST) + 1e-6).all(), (ours - ref_cost).max()


def _gradik_params(rng_arm, n, iterations=50):
    from av_aloha_b200 import capi
    p = capi.GradIKParams()
    p.step_size, p.min_cost_delta, p.max_iterations = 0.0001, 1.0e-12, iterations
    p.position_weight, p.rotation_weight = 500.0, 100.0
    p.position_threshold = p.rotation_threshold = 0.001
    p.max_pos_diff, p.max_rot_diff, p.joint_p = 0.1, 0.3, 0.9
    half = 0.5 * (rng_arm[:, 1] - rng_arm[:, 0])
    cw = [10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0]
    for k in range(7):
        p.joint_center_weight[k] = cw[k] / half[k] if k < n else 0.0
        p.joint_displacement_weight[k] = 50.0 if k < n else 0.0
    return p


# ------------------------------------------------------------------ CPU: kernel source through the emulator
@pytest.fixture(scope="module")
def emu_batch(slot_model_path):
    from tests.emu import emu
    return emu, emu.EmuBatch(slot_model_path, 1)


@pytest.mark.parametrize("arm", [0, 1, 2])
def test_emu_fk_matches_reference(gold, emu_batch, arm):
    emu, eb = emu_batch
    T = emu.emu_fk(eb, arm, gold[f"fk_q_{arm}"])
    assert np.abs(T - gold[f"fk_T_{arm}"]).max() <= TOL_FK


@pytest.mark.parametrize("arm", [0, 1, 2])
def test_emu_jacobian_matches_reference(gold, emu_batch, arm):
    """create_jac_fn (kinematics.py:28-52): golden = the reference's own jacobian() on the FK test poses"""
    emu, eb = emu_batch
    J = emu.emu_jac(eb, arm, gold[f"fk_q_{arm}"])
    assert J.shape == gold[f"jac_J_{arm}"].shape and np.abs(J - gold[f"jac_J_{arm}"]).max() <= TOL_FK


@pytest.mark.parametrize("arm", [0, 1, 2])
@pytest.mark.parametrize("tag", ["sim", "real"])
def test_emu_diffik_matches_reference(gold, emu_batch, arm, tag):
    emu, eb = emu_batch
    args = (gold[f"diffik_{tag}_q_{arm}"], gold[f"diffik_{tag}_pos_{arm}"], gold[f"diffik_{tag}_quat_{arm}"])
    _check_diffik(emu.emu_diffik(eb, arm, *args, _diffik_params(arm, tag)),
                  emu.emu_diffik(eb, arm, *args, _diffik_params(arm, tag, 1)), gold, arm, tag)


@pytest.mark.parametrize("arm", [0, 1, 2])
def test_emu_gradik_matches_reference(gold, emu_batch, arm, slot_model_path):
    from av_aloha_b200 import model_io
    emu, eb = emu_batch
    avm = model_io.load_avm(slot_model_path)
    n = int(avm["ik_ndof"][arm])
    args = (gold[f"gradik_q_{arm}"], gold[f"gradik_pos_{arm}"], gold[f"gradik_quat_{arm}"])
    _check_gradik(emu.emu_gradik(eb, arm, *args, _gradik_params(avm["ik_range"][arm], n)),
                  emu.emu_gradik(eb, arm, *args, _gradik_params(avm["ik_range"][arm], n, 8)), gold, arm, avm)


# ------------------------------------------------------------------ GPU: through the C-ABI
@pytest.fixture(scope="module")
def gpu_model(slot_model_path):
    from av_aloha_b200 import capi
    return capi.Model(slot_model_path, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
def test_gpu_fk(gold, gpu_model, arm):
    from av_aloha_b200 import kinematics
    T = kinematics.create_fk_fn(gpu_model, arm)(gold[f"fk_q_{arm}"])
    assert np.abs(T - gold[f"fk_T_{arm}"]).max() <= TOL_FK


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
def test_gpu_jacobian(gold, gpu_model, arm):
    from av_aloha_b200 import kinematics
    jac = kinematics.create_jac_fn(gpu_model, arm)
    J = jac(gold[f"fk_q_{arm}"])
    assert J.shape == gold[f"jac_J_{arm}"].shape and np.abs(J - gold[f"jac_J_{arm}"]).max() <= TOL_FK
    one = jac(gold[f"fk_q_{arm}"][0])                                    # single pose, like the reference's jacobian(theta)
    assert one.shape == gold[f"jac_J_{arm}"][0].shape and np.abs(one - gold[f"jac_J_{arm}"][0]).max() <= TOL_FK


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
@pytest.mark.parametrize("tag", ["sim", "real"])
def test_gpu_diffik(gold, gpu_model, arm, tag):
    from av_aloha_b200 import kinematics
    home = {0: [0, -0.082, 1.06, 0, -0.953, 0], 1: [0, -0.082, 1.06, 0, -0.953, 0], 2: [0, -0.8, 0.8, 0, 0.5, 0, 0]}[arm]
    kw = dict(kinematics.DIFFIK_SIM if tag == "sim" else kinematics.DIFFIK_REAL, q0=home)
    args = (gold[f"diffik_{tag}_q_{arm}"], gold[f"diffik_{tag}_pos_{arm}"], gold[f"diffik_{tag}_quat_{arm}"])
    ctl, ctl1 = kinematics.DiffIK(gpu_model, arm, **kw), kinematics.DiffIK(gpu_model, arm, **dict(kw, iterations=1))
    _check_diffik(ctl.run(*args), ctl1.run(*args), gold, arm, tag)
    one = ctl1.run(args[0][0], args[1][0], args[2][0])                   # single-problem call, like the reference's run()
    assert one.shape == (len(home),) and np.abs(one - gold[f"diffik_{tag}_out1_{arm}"][0]).max() <= TOL_DIFFIK1


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
def test_gpu_gradik(gold, gpu_model, arm):
    from av_aloha_b200 import kinematics
    n = (6, 6, 7)[arm]
    kw = dict(kinematics.GRADIK_SIM, joint_center_weight=(10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0)[:n],
              joint_displacement_weight=(50.0,) * n)
    args = (gold[f"gradik_q_{arm}"], gold[f"gradik_pos_{arm}"], gold[f"gradik_quat_{arm}"])
    ctl, ctl8 = kinematics.GradIK(gpu_model, arm, **kw), kinematics.GradIK(gpu_model, arm, **dict(kw, max_iterations=8))
    from av_aloha_b200 import model_io
    _check_gradik(ctl.run(*args), ctl8.run(*args), gold, arm, model_io.load_avm(gpu_model.avm_path))


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 2])
def test_gpu_gradik_group_and_thread_kernels_agree(gold, gpu_model, arm, monkeypatch):
    """The 16-lanes-per-problem kernel (small batches, low latency) and the thread-per-problem kernel (large batches) do the same
    evaluations with the same arithmetic: both forms against the golden vectors, and against each other."""
    from av_aloha_b200 import kinematics, model_io
    n = (6, 6, 7)[arm]
    kw = dict(kinematics.GRADIK_SIM, joint_center_weight=(10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0)[:n],
              joint_displacement_weight=(50.0,) * n)
    args = (gold[f"gradik_q_{arm}"], gold[f"gradik_pos_{arm}"], gold[f"gradik_quat_{arm}"])
    out = {}
    for form in ("0", "1"):
        monkeypatch.setenv("AVSIM_GRADIK_GROUP", form)
        ctl, ctl8 = kinematics.GradIK(gpu_model, arm, **kw), kinematics.GradIK(gpu_model, arm, **dict(kw, max_iterations=8))
        out[form] = (ctl.run(*args), ctl8.run(*args))
        _check_gradik(out[form][0], out[form][1], gold, arm, model_io.load_avm(gpu_model.avm_path))
    assert np.abs(out["0"][1] - out["1"][1]).max() <= 1e-6       # 8 iterations: same iterates (fp32 output)
    assert np.median(np.abs(out["0"][0] - out["1"][0])) <= 1e-6


# ------------------------------------------------------------------ create_safety_fn (reference kinematics.py:54-135)
SAFETY_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "safety_golden.npz")


def _check_safety(fn, sg, arm):
    """golden = (ok, message) of the reference's own safety_fn on the same inputs (tools/gen_safety_golden.py); the codes
    index the reference's messages in the order of its early returns"""
    from av_aloha_b200.kinematics import SAFETY_MESSAGES
    q, ctrl, T, has_T, code = (sg[f"{k}_{arm}"] for k in ("q", "ctrl", "T", "hasT", "code"))
    assert set(np.unique(code)) == set(range(6))                       # every return path of the reference is exercised
    ok_b, code_b = fn(q[has_T], ctrl[has_T], T[has_T])                  # batched, with a target pose
    assert np.array_equal(code_b, code[has_T]) and np.array_equal(ok_b, code[has_T] == 0)
    ok_n, code_n = fn(q[~has_T], ctrl[~has_T])                          # batched, Taction=None
    assert np.array_equal(code_n, code[~has_T])
    for i in (0, 1, 2, 3, 8, 9):                                        # single calls return the reference's (bool, message)
        ok, msg = fn(q[i], ctrl[i], T[i] if has_T[i] else None)
        assert (ok, msg) == (code[i] == 0, SAFETY_MESSAGES[code[i]])


@pytest.mark.parametrize("arm", [0, 1, 2])
def test_emu_safety_fn_matches_reference(emu_batch, arm, slot_model_path):
    from av_aloha_b200 import kinematics, model_io
    emu, eb = emu_batch
    sg, avm = np.load(SAFETY_GOLD), model_io.load_avm(slot_model_path)
    n = int(avm["ik_ndof"][arm])
    fk = lambda q: emu.emu_fk(eb, arm, np.atleast_2d(q))   # noqa: E731
    fn = kinematics._make_safety_fn(fk, avm["ik_range"][arm, :n], sg[f"bounds_{arm}"], 0.01, 1.0, 0.2, 3.0)
    _check_safety(fn, sg, arm)


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
def test_gpu_safety_fn_matches_reference(gpu_model, arm):
    from av_aloha_b200 import kinematics
    sg = np.load(SAFETY_GOLD)
    _check_safety(kinematics.create_safety_fn(gpu_model, arm, sg[f"bounds_{arm}"]), sg, arm)
