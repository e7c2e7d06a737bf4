"""IK parity against golden vectors produced by RUNNING the reference's own numba code (tools/gen_ik_golden.py).

CPU leg  : the CUDA IK kernels' source, compiled for the host by the warp-emulation harness (tests/emu), against the
           golden vectors -- checks the kernel arithmetic in the GPU-less container.
GPU leg  : the same kernels through the C-ABI (avsim_fk / avsim_diffik / avsim_gradik) on cuda:0.
Tolerance: 1e-5 abs for FK / DiffIK (the reference itself rounds through float32 in quat2mat,
           transform_utils.py:66; our I/O is fp32, arithmetic fp64), 2e-5 for GradIK (50 descent iterations).
"""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ik_golden.npz")
TOL_FK, TOL_DIFFIK, TOL_GRADIK = 1e-5, 1e-5, 2e-5


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _diffik_params(arm, tag):
    from av_aloha_b200 import capi
    home = {0: [0, -0.082, 1.06, 0, -0.953, 0, 0], 1: [0, -0.082, 1.06, 0, -0.953, 0, 0], 2: [0, -0.8, 0.8, 0, 0.5, 0, 0]}[arm]
    p = capi.DiffIKParams()
    p.k_pos = p.k_ori = 0.9 if tag == "sim" else 0.3
    p.integration_dt = 0.04 if tag == "sim" else 0.02
    p.damping, p.max_angvel, p.iterations = 1.0e-4, 3.14, 10
    for k, v in enumerate([10.0, 10.0, 10.0, 10.0, 5.0, 5.0, 5.0]):
        p.k_null[k] = v
        p.q0[k] = home[k]
    return p


def _gradik_params(rng_arm, n):
    from av_aloha_b200 import capi
    p = capi.GradIKParams()
    p.step_size, p.min_cost_delta, p.max_iterations = 0.0001, 1.0e-12, 50
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
@pytest.mark.parametrize("tag", ["sim", "real"])
def test_emu_diffik_matches_reference(gold, emu_batch, arm, tag):
    emu, eb = emu_batch
    out = emu.emu_diffik(eb, arm, gold[f"diffik_{tag}_q_{arm}"], gold[f"diffik_{tag}_pos_{arm}"],
                         gold[f"diffik_{tag}_quat_{arm}"], _diffik_params(arm, tag))
    assert np.abs(out - gold[f"diffik_{tag}_out_{arm}"]).max() <= TOL_DIFFIK


@pytest.mark.parametrize("arm", [0, 1, 2])
def test_emu_gradik_matches_reference(gold, emu_batch, arm, slot_model_path):
    from av_aloha_b200 import model_io
    emu, eb = emu_batch
    avm = model_io.load_avm(slot_model_path)
    n = int(avm["ik_ndof"][arm])
    out = emu.emu_gradik(eb, arm, gold[f"gradik_q_{arm}"], gold[f"gradik_pos_{arm}"], gold[f"gradik_quat_{arm}"],
                         _gradik_params(avm["ik_range"][arm], n))
    assert np.abs(out - gold[f"gradik_out_{arm}"]).max() <= TOL_GRADIK


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
@pytest.mark.parametrize("tag", ["sim", "real"])
def test_gpu_diffik(gold, gpu_model, arm, tag):
    from av_aloha_b200 import kinematics
    home = {0: [0, -0.082, 1.06, 0, -0.953, 0], 1: [0, -0.082, 1.06, 0, -0.953, 0], 2: [0, -0.8, 0.8, 0, 0.5, 0, 0]}[arm]
    kw = dict(kinematics.DIFFIK_SIM if tag == "sim" else kinematics.DIFFIK_REAL, q0=home)
    ctl = kinematics.DiffIK(gpu_model, arm, **kw)
    out = ctl.run(gold[f"diffik_{tag}_q_{arm}"], gold[f"diffik_{tag}_pos_{arm}"], gold[f"diffik_{tag}_quat_{arm}"])
    assert np.abs(out - gold[f"diffik_{tag}_out_{arm}"]).max() <= TOL_DIFFIK
    one = ctl.run(gold[f"diffik_{tag}_q_{arm}"][0], gold[f"diffik_{tag}_pos_{arm}"][0], gold[f"diffik_{tag}_quat_{arm}"][0])
    assert one.shape == (len(home),) and np.abs(one - gold[f"diffik_{tag}_out_{arm}"][0]).max() <= TOL_DIFFIK


@pytest.mark.gpu
@pytest.mark.parametrize("arm", [0, 1, 2])
def test_gpu_gradik(gold, gpu_model, arm):
    from av_aloha_b200 import kinematics
    n = (6, 6, 7)[arm]
    kw = dict(kinematics.GRADIK_SIM, joint_center_weight=(10.0, 10.0, 1.0, 50.0, 1.0, 1.0, 1.0)[:n],
              joint_displacement_weight=(50.0,) * n)
    ctl = kinematics.GradIK(gpu_model, arm, **kw)
    out = ctl.run(gold[f"gradik_q_{arm}"], gold[f"gradik_pos_{arm}"], gold[f"gradik_quat_{arm}"])
    assert np.abs(out - gold[f"gradik_out_{arm}"]).max() <= TOL_GRADIK
