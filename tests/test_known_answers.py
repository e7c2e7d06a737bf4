"""Known-answer pins that do not go through the restatement: closed forms of the integrator MuJoCo documents
(semi-implicit Euler: v += h a, then q += h v_new; free-joint quaternions advanced by the exponential map of h w).

The physics oracle is "parity unpinned" against MuJoCo itself (MuJoCo is not installable here and the
reference tree holds no golden trajectories).  What CAN be pinned without MuJoCo is every place where the pipeline has an
analytic answer.  Here: a free body (the SlotInsertion stick, lifted 0.5 m above the table, spinning about its long axis)
falls for N substeps of h = 2 ms without touching anything:

    z_N  = z_0 - g h^2 N (N + 1) / 2          vz_N = -g h N
    quat = (cos(w h N / 2), 0, 0, sin(w h N / 2)),  angular velocity unchanged (spin about a principal axis: w x I w = 0)

checked for the fp64 oracle (1e-12), for the CUDA kernel source run by the host emulator (fp32: 1e-5) and, on a GPU, for the
kernels through the C-ABI (1e-5).  The robot arms keep moving under their PD controllers meanwhile; they do not matter here.
"""
import numpy as np
import pytest

HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0])
FP = np.array([[0.0, 0.12, 0.0], [0.02, -0.05, 0.5]])       # slot on the table, stick 0.5 m up
N, W_SPIN = 50, 3.0


def _layout(path):
    from av_aloha_b200 import model_io
    avm = model_io.load_avm(path)
    qadr = int(avm["free_qadr"][1])                          # stick_joint: qpos [x y z qw qx qy qz]
    return avm, qadr


def _expected(avm):
    h, g = float(avm["timestep"][0]), -float(avm["gravity"][2])
    th = W_SPIN * h * N
    return FP[1, 2] - g * h * h * N * (N + 1) / 2, -g * h * N, np.array([np.cos(th / 2), 0.0, 0.0, np.sin(th / 2)])


def test_oracle_free_fall_and_spin_closed_form(slot_model_path):
    from oracle.oracle import OracleEnv, OracleModel
    avm, qadr = _layout(slot_model_path)
    o = OracleEnv(OracleModel(slot_model_path))
    o.reset(free_pos=FP)
    dof = o.model.nv - 6                                      # the stick's free joint holds the last six dofs
    assert qadr - (o.model.nq - o.model.nv) + 1 == dof       # two free joints before/at it: one quaternion slot each
    o.qvel[dof + 5] = W_SPIN                                  # local z = the stick's long axis
    for _ in range(N):
        o.substep()
    z, vz, quat = _expected(avm)
    assert abs(o.qpos[qadr + 2] - z) <= 1e-12 and abs(o.qvel[dof + 2] - vz) <= 1e-12
    assert np.abs(o.qpos[qadr + 3:qadr + 7] - quat).max() <= 1e-12
    assert np.abs(o.qvel[dof + 3:dof + 6] - [0, 0, W_SPIN]).max() <= 1e-12
    assert np.abs(o.qpos[qadr:qadr + 2] - FP[1, :2]).max() <= 1e-12      # no lateral drift


def test_kernel_source_free_fall_and_spin_closed_form(slot_model_path):
    """the CUDA step kernel's source, compiled for the host by the warp emulator (fp32)"""
    from tests.emu.emu import EmuBatch
    avm, qadr = _layout(slot_model_path)
    eb = EmuBatch(slot_model_path, 1)
    eb.set_options(8)
    eb.reset(FP[None])
    dof = eb.nv - 6
    eb.qvel[0, dof + 5] = W_SPIN
    act = HOME.copy(); act[6] = act[13] = 1.0
    eb.step(act[None].astype(np.float32), N)
    z, vz, quat = _expected(avm)
    assert eb.status[0] == 0
    assert abs(eb.qpos[0, qadr + 2] - z) <= 1e-5 and abs(eb.qvel[0, dof + 2] - vz) <= 1e-5
    assert np.abs(eb.qpos[0, qadr + 3:qadr + 7] - quat).max() <= 1e-5
    assert np.abs(eb.qvel[0, dof + 3:dof + 6] - [0, 0, W_SPIN]).max() <= 1e-5


@pytest.mark.gpu
def test_gpu_free_fall_and_spin_closed_form(slot_model_path):
    import torch
    from av_aloha_b200 import capi
    avm, qadr = _layout(slot_model_path)
    model = capi.Model(slot_model_path, 0)
    B = 3
    b = capi.Batch(model, B, seed=0)
    b.set_options(solver_iters=8)
    b.reset(free_pos=np.tile(FP[None], (B, 1, 1)))
    dof = model.nv - 6
    qvel = b.get(capi.QVEL)
    qvel[:, dof + 5] = W_SPIN
    b.set(capi.QVEL, qvel)
    act = HOME.copy(); act[6] = act[13] = 1.0
    b.step(torch.as_tensor(np.tile(act, (B, 1)), dtype=torch.float32, device="cuda"), N)
    qpos, qv = b.get(capi.QPOS).cpu().numpy().astype(np.float64), b.get(capi.QVEL).cpu().numpy().astype(np.float64)
    z, vz, quat = _expected(avm)
    assert int(b.get(capi.STATUS).max().item()) == 0
    assert np.abs(qpos[:, qadr + 2] - z).max() <= 1e-5 and np.abs(qv[:, dof + 2] - vz).max() <= 1e-5
    assert np.abs(qpos[:, qadr + 3:qadr + 7] - quat).max() <= 1e-5
    assert np.abs(qv[:, dof + 3:dof + 6] - [0, 0, W_SPIN]).max() <= 1e-5
    b.close()
