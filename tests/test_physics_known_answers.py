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
    eb.set_solver("pgs"); eb.set_options(8)
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
    b.set_solver("pgs"); b.set_options(solver_iters=8)
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


# ---------------------------------------------------------------------------------------------------------------------
# Static equilibrium (Newton's first law through the whole contact pipeline): the stick (0.3536 kg, 4 box-box contacts
# with the table) and the slot (100 kg) at rest on the table.  Once settled, the constraint solve must return contact
# forces whose resultant cancels gravity: qfrc_constraint_z = m g, qacc = 0 -- independent of solref / solimp / cone
# details, which only set the penetration at which that happens.
REST = np.array([[0.0, 0.12, 0.0], [0.02, -0.05, 0.0]])


def _settled_oracle(path):
    from oracle.oracle import OracleEnv, OracleModel
    o = OracleEnv(OracleModel(path))
    o.set_solver("pgs"); o.set_options(max_iter=200, tol=1e-12, warmstart=1)
    o.reset(free_pos=REST)
    act = HOME.copy(); act[6] = act[13] = 1.0
    for _ in range(25):                                       # 1 s: the soft contacts (solref 0.02 s) have long settled
        o.step(act)
    o.forward()
    return o


def test_oracle_resting_bodies_carry_their_weight(slot_model_path):
    o = _settled_oracle(slot_model_path)
    g = 9.81
    for dof in (o.model.nv - 6, o.model.nv - 12):             # stick, slot (free joints: the last twelve dofs)
        m = o.M[dof + 2, dof + 2]
        assert abs(o.qfrc_constraint[dof + 2] / (m * g) - 1.0) <= 1e-9
        assert np.abs(o.qacc[dof:dof + 6]).max() <= 1e-9 and np.abs(o.qvel[dof:dof + 6]).max() <= 1e-9
    assert o.M[o.model.nv - 6 + 2, o.model.nv - 6 + 2] == pytest.approx(0.3536, rel=1e-6)       # task_slot_insertion.xml stick mass
    assert o.ncon >= 8


def test_kernel_source_resting_bodies_carry_their_weight(slot_model_path):
    """the CUDA forward pass (emulated, fp32) at the settled state: gravity is cancelled to 1e-4 of g"""
    from tests.emu.emu import EmuBatch
    o = _settled_oracle(slot_model_path)
    eb = EmuBatch(slot_model_path, 1)
    eb.set_solver("pgs"); eb.set_options(50)
    eb.reset(REST[None])
    eb.qpos[0, :], eb.qvel[0, :], eb.ctrl[0, :] = o.qpos.astype(np.float32), 0.0, o.ctrl.astype(np.float32)
    eb.forward()
    assert eb.ncon[0] == o.ncon and eb.status[0] == 0
    for dof in (eb.nv - 6, eb.nv - 12):
        assert abs(eb.qacc_smooth[0, dof + 2] + 9.81) <= 1e-4            # free body: smooth acceleration = gravity
        assert abs(eb.qacc[0, dof + 2]) <= 1e-3                          # ... cancelled by the contact forces


@pytest.mark.gpu
def test_gpu_resting_bodies_carry_their_weight(slot_model_path):
    from av_aloha_b200 import capi
    o = _settled_oracle(slot_model_path)
    model = capi.Model(slot_model_path, 0)
    b = capi.Batch(model, 2, seed=0)
    b.set_solver("pgs"); b.set_options(solver_iters=50)
    b.reset(free_pos=np.tile(REST[None], (2, 1, 1)))
    b.set(capi.QPOS, np.tile(o.qpos.astype(np.float32), (2, 1)))
    b.set(capi.QVEL, np.zeros((2, model.nv), np.float32))
    b.set(capi.CTRL, np.tile(o.ctrl.astype(np.float32), (2, 1)))
    b.forward()
    qacc, smooth = b.get(capi.QACC).cpu().numpy(), b.get(capi.QACC_SMOOTH).cpu().numpy()
    assert (b.get(capi.NCON).cpu().numpy() == o.ncon).all() and int(b.get(capi.STATUS).max().item()) == 0
    for dof in (model.nv - 6, model.nv - 12):
        assert np.abs(smooth[:, dof + 2] + 9.81).max() <= 1e-4
        assert np.abs(qacc[:, dof + 2]).max() <= 1e-3
    b.close()


# ---------------------------------------------------------------------------------------------------------------------
# Rigid-body inertia of the task objects from the MJCF numbers alone (reference assets/task_slot_insertion.xml:5-16):
# box of half sizes (a, b, c) and mass m: I_com = m/3 diag(b^2 + c^2, a^2 + c^2, a^2 + b^2), shifted to the body origin by the
# parallel-axis term m (|r|^2 1 - r r^T); a free joint's mass block is [[m 1, -m [r]x], [m [r]x, I_origin]].  Default
# density 1000 kg/m^3 for geoms without a mass attribute.  Pins the model compiler (mesh-free part) and the CRB stage.
def _box(m, half, r):
    a, b, c = half
    I = m / 3.0 * np.diag([b * b + c * c, a * a + c * c, a * a + b * b])
    r = np.asarray(r, float)
    return I + m * (r @ r * np.eye(3) - np.outer(r, r)), m * r


def _free_block(parts):
    M = np.zeros((6, 6))
    for m, half, r in parts:
        I, mr = _box(m, half, r)
        M[:3, :3] += m * np.eye(3)
        M[3:, 3:] += I
        K = np.array([[0, -mr[2], mr[1]], [mr[2], 0, -mr[0]], [-mr[1], mr[0], 0]])      # m [r]x
        M[:3, 3:] -= K
        M[3:, :3] += K
    return M


STICK = [(1000.0 * 8 * 0.17 * 0.013 * 0.02, (0.17, 0.013, 0.02), (0, 0, 0.02))]
SLOT = [(50.0, (0.1, 0.015, 0.02), (0, 0.032, 0.02)), (50.0, (0.1, 0.015, 0.02), (0, -0.032, 0.02))]


def test_free_body_mass_blocks_match_the_mjcf_numbers(slot_model_path):
    from oracle.oracle import OracleEnv, OracleModel
    from tests.emu.emu import EmuBatch
    o = OracleEnv(OracleModel(slot_model_path))
    o.reset(free_pos=REST)                                    # identity orientation: body frame = world frame
    o.forward()
    nv = o.model.nv
    eb = EmuBatch(slot_model_path, 1)
    eb.reset(REST[None])
    eb.forward()
    for dof, parts in ((nv - 6, STICK), (nv - 12, SLOT)):
        want = _free_block(parts)
        got = o.M[dof:dof + 6, dof:dof + 6]
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max())
        assert np.abs(eb.mass_diag[0, dof:dof + 6] - np.diag(want)).max() <= 1e-6 * np.abs(np.diag(want)).max()


# ---------------------------------------------------------------------------------------------------------------------
# Principle of virtual work: at zero velocity the bias force of the recursive Newton-Euler pass is the gradient of the
# potential energy U(q) = sum_b m_b g z_com,b(q).  U comes from the KINEMATICS stage (body poses) and the model's masses /
# centre-of-mass offsets; the bias from the RNE stage -- two separate code paths that only agree when the spatial
# transforms, the joint axes (cdof) and the inertial frames are all right.  23 hinge / slide joints of the three arms.
def _quat_rot(q, v):
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return R @ v


def _potential_gradient(o, avm, q0, nj, d=1e-6):
    def U(q):
        o.qpos[:] = q
        o.qvel[:] = 0
        o.forward()
        return sum(avm["body_mass"][b] * 9.81 * (o.xpos[b] + _quat_rot(o.xquat[b], avm["body_ipos"][b]))[2]
                   for b in range(1, o.model.nbody))
    g = np.zeros(nj)
    for i in range(nj):
        qp, qm = q0.copy(), q0.copy()
        qp[i] += d
        qm[i] -= d
        g[i] = (U(qp) - U(qm)) / (2 * d)
    return g


@pytest.fixture(scope="module")
def virtual_work(slot_model_path):
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleEnv, OracleModel
    avm = model_io.load_avm(slot_model_path)
    o = OracleEnv(OracleModel(slot_model_path))
    o.reset(free_pos=np.array([[0, 0.12, 0.3], [0.02, -0.05, 0.5]]))
    nj = 23                                                   # hinge / slide joints: qpos index == dof index
    assert (avm["jnt_type"][:nj] != 0).all() and (avm["jnt_qposadr"][:nj] == np.arange(nj)).all()
    q0 = o.qpos.copy()
    q0[:nj] += np.random.default_rng(0).normal(0, 0.2, nj)
    grad = _potential_gradient(o, avm, q0, nj)
    o.qpos[:] = q0
    o.qvel[:] = 0
    o.forward()
    return q0, grad, o.qfrc_bias[:nj].copy(), o.ctrl.copy(), nj


def test_oracle_bias_force_is_the_potential_gradient(virtual_work):
    q0, grad, bias, ctrl, nj = virtual_work
    assert np.abs(bias).max() > 1.0                           # shoulders carry a few N m
    assert np.abs(bias - grad).max() <= 1e-6


def test_kernel_source_bias_force_is_the_potential_gradient(virtual_work, slot_model_path):
    from tests.emu.emu import EmuBatch
    q0, grad, bias, ctrl, nj = virtual_work
    eb = EmuBatch(slot_model_path, 1)
    eb.reset(REST[None])
    eb.qpos[0, :], eb.qvel[0, :], eb.ctrl[0, :] = q0.astype(np.float32), 0.0, ctrl.astype(np.float32)
    eb.forward()
    assert np.abs(eb.qfrc_bias[0, :nj] - grad).max() <= 1e-4 * np.abs(grad).max()


@pytest.mark.gpu
def test_gpu_bias_force_is_the_potential_gradient(virtual_work, slot_model_path):
    from av_aloha_b200 import capi
    q0, grad, bias, ctrl, nj = virtual_work
    model = capi.Model(slot_model_path, 0)
    b = capi.Batch(model, 1, seed=0)
    b.reset(free_pos=REST[None])
    b.set(capi.QPOS, q0.astype(np.float32)[None])
    b.set(capi.QVEL, np.zeros((1, model.nv), np.float32))
    b.forward()
    got = b.get(capi.QFRC_BIAS).cpu().numpy()[0, :nj]
    assert np.abs(got - grad).max() <= 1e-4 * np.abs(grad).max()
    b.close()


# ---------------------------------------------------------------------------------------------------------------------
# Kinetic energy: (1/2) v^T M v must equal the sum over bodies of (1/2) m |v_com|^2 + (1/2) w^T I w (+ the joints' armature),
# with v_com and w obtained by finite-differencing the KINEMATICS stage's body poses along v.  Pins the composite-rigid-body
# mass matrix (and the model's inertial frames) against an independent route.  Unit vectors give the diagonal (also
# checked for the fp32 kernels); random directions give the full quadratic form.
def _quat_mat(q):
    return np.stack([_quat_rot(q, e) for e in np.eye(3)], axis=1)


def _kinetic_energy_fd(o, avm, q0, v, nj, d=1e-6):
    def poses(q):
        o.qpos[:] = q
        o.qvel[:] = 0
        o.forward()
        R = [_quat_mat(o.xquat[b]) for b in range(o.model.nbody)]
        return [o.xpos[b] + R[b] @ avm["body_ipos"][b] for b in range(o.model.nbody)], R
    qp, qm = q0.copy(), q0.copy()
    qp[:nj] += d * v
    qm[:nj] -= d * v
    (cp, Rp), (cm, Rm) = poses(qp), poses(qm)
    T = 0.5 * float(np.sum(avm["dof_armature"][:nj] * v * v))
    for b in range(1, o.model.nbody):
        m = avm["body_mass"][b]
        if m == 0:
            continue
        vc = (cp[b] - cm[b]) / (2 * d)
        S = (Rp[b] @ Rm[b].T - Rm[b] @ Rp[b].T) / (4 * d)      # [w]x to first order
        w = np.array([S[2, 1], S[0, 2], S[1, 0]])
        i6 = avm["body_inertia"][b]
        Ib = np.array([[i6[0], i6[3], i6[4]], [i6[3], i6[1], i6[5]], [i6[4], i6[5], i6[2]]])
        R0 = _quat_mat(_mid_quat(o, q0, b))
        T += 0.5 * m * vc @ vc + 0.5 * w @ (R0 @ Ib @ R0.T) @ w
    return T


def _mid_quat(o, q0, b):
    o.qpos[:] = q0
    o.qvel[:] = 0
    o.forward()
    return o.xquat[b].copy()


def test_mass_matrix_reproduces_the_kinetic_energy(virtual_work, slot_model_path):
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleEnv, OracleModel
    from tests.emu.emu import EmuBatch
    q0, _, _, ctrl, nj = virtual_work
    avm = model_io.load_avm(slot_model_path)
    o = OracleEnv(OracleModel(slot_model_path))
    o.reset(free_pos=np.array([[0, 0.12, 0.3], [0.02, -0.05, 0.5]]))
    o.qpos[:] = q0
    o.forward()
    M = o.M[:nj, :nj].copy()
    diag_fd = np.array([2 * _kinetic_energy_fd(o, avm, q0, np.eye(nj)[i], nj) for i in range(nj)])
    assert np.abs(np.diag(M) - diag_fd).max() <= 1e-6 * np.abs(diag_fd).max()
    rng = np.random.default_rng(1)
    for _ in range(4):
        v = rng.normal(0, 1, nj)
        T = _kinetic_energy_fd(o, avm, q0, v, nj)
        assert abs(0.5 * v @ M @ v - T) <= 1e-6 * T
    eb = EmuBatch(slot_model_path, 1)                         # the CUDA CRB stage (emulated, fp32): diagonal of M
    eb.reset(REST[None])
    eb.qpos[0, :], eb.qvel[0, :] = q0.astype(np.float32), 0.0
    eb.forward()
    assert np.abs(eb.mass_diag[0, :nj] - diag_fd).max() <= 1e-5 * np.abs(diag_fd).max()


# ---------------------------------------------------------------------------------------------------------------------
# Feasibility of the contact solve in a contact-rich state (the scripted grasp of the bench workload, 20 contacts, all condim 6
# with elliptic cones, impratio 100): every normal force is non-negative and every friction force lies inside its elliptic cone
# sum_i (f_i / mu_i)^2 <= f_n^2 -- a property of the solution the constraint solver must return whatever its path to it.
def test_oracle_contact_forces_are_cone_feasible(slot_model_path):
    from av_aloha_b200 import workload
    from oracle.oracle import OracleEnv, OracleModel
    obj = workload.sample_object_positions(5, 11)
    acts = workload.slot_insertion_script(300, obj, 11)
    o = OracleEnv(OracleModel(slot_model_path))
    o.set_solver("pgs"); o.set_options(max_iter=100, tol=1e-10, warmstart=1)
    o.reset(free_pos=obj[0])
    for t in range(215):
        o.step(acts[t, 0].astype(np.float64))
    o.forward()
    C, f = o.contacts(), o.efc_force
    live = [c for c in C if not int(c[16])]
    r = o.nefc - sum(int(c[15]) for c in live)                 # contact rows follow the scalar rows
    active = 0
    for c in live:
        d = int(c[15])
        fc, mu = f[r:r + d], c[17:17 + d - 1]
        assert fc[0] >= -1e-12
        assert np.sqrt(np.sum((fc[1:] / mu) ** 2)) <= fc[0] * (1 + 1e-9) + 1e-12
        active += fc[0] > 1e-9
        r += d
    assert o.ncon >= 12 and active >= 8 and o.reward >= 1     # a real grasp state, not an empty one


# ---------------------------------------------------------------------------------------------------------------------
# Narrowphase closed forms (MuJoCo's conventions: dist = signed gap, normal from geom1 to geom2, position = midpoint between the
# two surfaces).  Table top at z = top; the stick box (half sizes 0.17 x 0.013 x 0.02) pushed `pen` into it gives four contacts at
# its bottom corners; a finger-pad sphere (r = 0.6 mm) gives one below its centre.  The sphere case caught a sign error in
# round 1: the position was |dist| / 2 ABOVE the box surface instead of below it.
def test_narrowphase_closed_forms(slot_model_path):
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleModel
    avm = model_io.load_avm(slot_model_path)
    names = model_io.load_names("slot_insertion", 3)["geom"]
    om = OracleModel(slot_model_path)
    gt, gs, gp = names.index("table"), names.index("stick"), names.index("left_left_g0")
    tpos, eye = avm["geom_pos"][gt], np.eye(3)
    top = tpos[2] + avm["geom_size"][gt][2]
    pen, half = 1e-3, avm["geom_size"][gs]
    for yaw in (0.0, 0.5):
        c, s = np.cos(yaw), np.sin(yaw)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
        centre = np.array([0.05, -0.03, top + half[2] - pen])
        out = om.collide_pair(gt, tpos, eye, gs, centre, Rz)
        assert len(out) == 4
        corners = np.array([centre + Rz @ (np.array([sx, sy, -1.0]) * half) for sx in (1, -1) for sy in (1, -1)])
        for row in out:
            assert abs(row[0] + pen) <= 1e-12 and np.abs(row[4:7] - [0, 0, 1]).max() <= 1e-12      # dist, normal table -> stick
            assert abs(row[3] - (top - pen / 2)) <= 1e-12                                           # midpoint height
            assert np.abs(corners[:, :2] - row[1:3]).sum(axis=1).min() <= 1e-12                     # at a bottom corner
        assert len({tuple(np.round(r[1:3], 9)) for r in out}) == 4
    r, pen = float(avm["geom_size"][gp][0]), 2e-4
    assert avm["geom_type"][gp] == 2 and r == pytest.approx(6e-4)
    for g1, g2, sign in ((gt, gp, 1.0), (gp, gt, -1.0)):                                            # both argument orders
        sc = np.array([0.1, 0.1, top + r - pen])
        args = (g1, tpos, eye, g2, sc, eye) if g1 == gt else (g1, sc, eye, g2, tpos, eye)
        out = om.collide_pair(*args)
        assert len(out) == 1
        assert abs(out[0][0] + pen) <= 1e-12 and np.abs(out[0][4:7] - [0, 0, sign]).max() <= 1e-12
        assert np.abs(out[0][1:4] - [0.1, 0.1, top - pen / 2]).max() <= 1e-12                        # midpoint, BELOW the table top
    assert len(om.collide_pair(gt, tpos, eye, gp, np.array([0.1, 0.1, top + r + 1e-6]), eye)) == 0  # separated: no contact


def test_kernel_source_contact_records_match_the_oracle_in_a_grasp_state(slot_model_path):
    """19 contacts incl. five finger-pad spheres on the stick and two pad-pad pairs: the CUDA narrowphase (emulated, fp32) returns
    the oracle's contact list -- same geoms in the same order, gaps and positions to 1e-6 m, normals to 1e-3."""
    from av_aloha_b200 import model_io, workload
    from oracle.oracle import OracleEnv, OracleModel
    from tests.emu.emu import EmuBatch
    avm = model_io.load_avm(slot_model_path)
    obj = workload.sample_object_positions(5, 11)
    acts = workload.slot_insertion_script(300, obj, 11)
    o = OracleEnv(OracleModel(slot_model_path))
    o.set_solver("pgs"); o.set_options(max_iter=8, tol=0.0, warmstart=2)
    o.reset(free_pos=obj[0])
    for t in range(215):
        o.step(acts[t, 0].astype(np.float64))
    o.forward()
    C = o.contacts()
    spheres = [c for c in C if 2 in (avm["geom_type"][int(c[13])], avm["geom_type"][int(c[14])])]
    assert len(C) >= 15 and len(spheres) >= 5
    eb = EmuBatch(slot_model_path, 1)
    eb.set_solver("pgs"); eb.set_options(8)
    eb.reset(obj[0][None])
    eb.qpos[0, :], eb.qvel[0, :], eb.ctrl[0, :] = o.qpos.astype(np.float32), o.qvel.astype(np.float32), o.ctrl.astype(np.float32)
    eb.forward()
    assert eb.ncon[0] == len(C)
    ec = eb.contacts[0].reshape(64, 16)[: len(C)]
    assert (ec[:, 7:9] == C[:, 13:15]).all()
    assert np.abs(ec[:, 0] - C[:, 0]).max() <= 1e-6 and np.abs(ec[:, 1:4] - C[:, 1:4]).max() <= 1e-6
    assert np.abs(ec[:, 4:7] - C[:, 4:7]).max() <= 1e-3


def test_mpr_closed_form_cylinder_on_table():
    """The convex (MPR) path on a shape with a closed form: the HookPackage hook cylinder (r = 6 mm, half height 0.1) standing on
    the table, pushed `pen` into it -> gap -pen, normal +z, midpoint position; with multiccd four more points on the rim
    (the two shapes are counter-rotated by 1e-3 rad about the first contact point: depth changes by at most 2e-3 x the lever, <= 2 r); lying on its side: a line contact, multiccd finds both ends."""
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleModel
    path = model_io.model_path("hook_package", 2)
    avm, names = model_io.load_avm(path), model_io.load_names("hook_package", 2)["geom"]
    om = OracleModel(path)
    gt, gh = names.index("table"), names.index("hook")
    assert avm["geom_type"][gh] == 5
    r, hh = (float(x) for x in avm["geom_size"][gh][:2])
    tpos, eye, pen = avm["geom_pos"][gt], np.eye(3), 1e-3
    top = tpos[2] + avm["geom_size"][gt][2]
    centre = np.array([0.1, 0.05, top + hh - pen])
    one = om.collide_pair(gt, tpos, eye, gh, centre, eye, multiccd=0)
    assert len(one) == 1
    assert abs(one[0][0] + pen) <= 1e-6 and np.abs(one[0][4:7] - [0, 0, 1]).max() <= 1e-6
    assert abs(one[0][3] - (top - pen / 2)) <= 1e-6 and np.linalg.norm(one[0][1:3] - centre[:2]) <= r + 1e-6
    five = om.collide_pair(gt, tpos, eye, gh, centre, eye, multiccd=1)
    assert len(five) == 5 and np.array_equal(five[0], one[0])
    for row in five[1:]:
        assert abs(np.linalg.norm(row[1:3] - centre[:2]) - r) <= 1e-4                 # on the rim
        assert abs(row[0] + pen) <= 2e-3 * 2 * r + 1e-6 and np.abs(row[4:7] - [0, 0, 1]).max() <= 1e-6
    Rx = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.0]])                                # axis along world -y
    side = om.collide_pair(gt, tpos, eye, gh, np.array([0.1, 0.05, top + r - pen]), Rx, multiccd=1)
    assert len(side) >= 3 and abs(side[0][0] + pen) <= 1e-6
    ys = sorted(row[2] - 0.05 for row in side)
    assert ys[0] <= -0.9 * hh and ys[-1] >= 0.9 * hh                                     # both ends of the line contact
    for row in side:
        assert abs(row[1] - 0.1) <= 1e-4 and np.abs(row[4:7] - [0, 0, 1]).max() <= 1e-5 and abs(row[0] + pen) <= 2e-3 * 2 * hh + 1e-6


def test_narrowphase_closed_form_sphere_sphere(slot_model_path):
    """two finger-pad spheres (r = 0.6 mm) overlapping by `pen` along an oblique axis: gap -pen, normal along the centre line from
    geom1 to geom2, position at the midpoint of the overlap"""
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleModel
    avm, names = model_io.load_avm(slot_model_path), model_io.load_names("slot_insertion", 3)["geom"]
    om = OracleModel(slot_model_path)
    g1, g2 = names.index("left_left_g0"), names.index("left_left_g1")
    r1, r2, pen, eye = float(avm["geom_size"][g1][0]), float(avm["geom_size"][g2][0]), 1e-4, np.eye(3)
    n = np.array([1.0, 2.0, -2.0]) / 3.0
    p1 = np.array([0.2, -0.1, 0.3])
    p2 = p1 + n * (r1 + r2 - pen)
    out = om.collide_pair(g1, p1, eye, g2, p2, eye)
    assert len(out) == 1
    assert abs(out[0][0] + pen) <= 1e-12 and np.abs(out[0][4:7] - n).max() <= 1e-12
    assert np.abs(out[0][1:4] - (p1 + n * (r1 - pen / 2))).max() <= 1e-12
    assert len(om.collide_pair(g1, p1, eye, g2, p1 + n * (r1 + r2 + 1e-7), eye)) == 0


def test_narrowphase_closed_form_tilted_box(slot_model_path):
    """the stick box rolled / pitched / both over the table: an edge gives two contacts at its ends, a corner gives one; gap = the
    depth of the lowest corners, normal +z, position at the midpoint between that corner and the table top"""
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleModel
    avm, names = model_io.load_avm(slot_model_path), model_io.load_names("slot_insertion", 3)["geom"]
    om = OracleModel(slot_model_path)
    gt, gs = names.index("table"), names.index("stick")
    tpos, eye, half, pen = avm["geom_pos"][gt], np.eye(3), avm["geom_size"][gs], 1e-3
    top = tpos[2] + avm["geom_size"][gt][2]

    def rot(axis, deg):
        c, s = np.cos(np.deg2rad(deg)), np.sin(np.deg2rad(deg))
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]]) if axis == "x" else np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])

    for R, npts in ((rot("x", 30), 2), (rot("y", 10), 2), (rot("x", 30) @ rot("y", 10), 1)):
        corners = np.array([R @ (np.array([sx, sy, sz]) * half) for sx in (1, -1) for sy in (1, -1) for sz in (1, -1)])
        low = -corners[:, 2].min()
        centre = np.array([0.05, -0.03, top + low - pen])
        lowest = centre + corners[np.abs(corners[:, 2] + low) <= 1e-12]          # the corners that touch first
        assert len(lowest) == npts
        out = om.collide_pair(gt, tpos, eye, gs, centre, R)
        assert len(out) == npts
        for row in out:
            assert abs(row[0] + pen) <= 1e-9 and np.abs(row[4:7] - [0, 0, 1]).max() <= 1e-9 and abs(row[3] - (top - pen / 2)) <= 1e-9
            assert np.abs(lowest[:, :2] - row[1:3]).sum(axis=1).min() <= 1e-9


# ---------------------------------------------------------------------------------------------------------------------
# Position actuators (reference assets/joint_position_actuators.xml + aloha_sim.xml:33-48, 95): force =
# clip(kp (clip(ctrl, ctrlrange) - q) - kv qdot, actuatorfrcrange).  Commands far beyond the pose saturate first the ctrl range
# (wrist_rotate: 10.4 x 3.14158 = 32.67 N m < 35; gripper: 0.037 is the open pose, force 0) and then the joint's force range
# (35 / 144 / 59 / 22 N m).  The oracle exposes qfrc_actuator; the CUDA smooth stage is checked through qacc_smooth at that state.
@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_actuator_force_saturation_closed_form(slot_model_path, sign):
    from av_aloha_b200 import model_io
    from oracle.oracle import OracleEnv, OracleModel
    from tests.emu.emu import EmuBatch
    avm = model_io.load_avm(slot_model_path)
    o = OracleEnv(OracleModel(slot_model_path))
    o.reset(free_pos=REST)
    rng = np.random.default_rng(3)
    o.qvel[:23] = rng.normal(0, 0.3, 23)
    o.ctrl[:] = o.ctrl + sign * 10.0
    ctrl_cmd = o.ctrl.copy()
    o.forward()
    kp, kv, qa, da = avm["act_kp"], avm["act_kv"], avm["act_qadr"], avm["act_dof"]
    u = np.clip(ctrl_cmd, avm["act_ctrl_lo"], avm["act_ctrl_hi"])
    want = np.zeros(o.model.nv)
    want[da] = kp * (u - o.qpos[qa]) - kv * o.qvel[da]
    lim = avm["dof_frc_limited"].astype(bool)
    want[lim] = np.clip(want[lim], avm["dof_frc_lo"][lim], avm["dof_frc_hi"][lim])
    assert np.abs(o.qfrc_actuator - want).max() <= 1e-9
    sat = np.isclose(np.abs(want[da]), avm["dof_frc_hi"][da])
    assert sat.sum() >= 12 and abs(abs(want[da[5]]) - 10.4 * 3.14158) <= 0.3 * 10.4 + 1e-9      # force range and ctrl range both bite
    eb = EmuBatch(slot_model_path, 1)
    eb.reset(REST[None])
    eb.qvel[0, :], eb.ctrl[0, :] = o.qvel.astype(np.float32), ctrl_cmd.astype(np.float32)
    eb.forward()
    scale = np.abs(o.qacc_smooth).max()
    assert np.abs(eb.qacc_smooth[0] - o.qacc_smooth).max() <= 1e-4 * scale


def test_oracle_friction_loss_is_dry_friction_and_fingers_stay_coupled(slot_model_path):
    """Scalar constraint rows during the reach of the scripted policy: the two finger-coupling equalities (J = [1, -1]) keep each
    gripper's fingers together, and the six friction-loss rows (shoulder 2.0, elbow 1.15 N m, aloha_sim.xml:40,44) behave as dry
    friction: |f| <= frictionloss always, and f = -frictionloss sign(qdot) on every joint that is moving."""
    from av_aloha_b200 import model_io, workload
    from oracle.oracle import OracleEnv, OracleModel
    avm = model_io.load_avm(slot_model_path)
    obj = workload.sample_object_positions(5, 11)
    acts = workload.slot_insertion_script(300, obj, 11)
    o = OracleEnv(OracleModel(slot_model_path))
    o.set_solver("pgs"); o.set_options(max_iter=100, tol=1e-10, warmstart=1)
    o.reset(free_pos=obj[0])
    for k in range(60):
        o.step(acts[k, 0].astype(np.float64))
    o.forward()
    nsc = o.nefc - sum(int(c[15]) for c in o.contacts() if not int(c[16]))
    J, f = o.efc_J[:nsc], o.efc_force[:nsc]
    fl_dofs = np.nonzero(avm["dof_frictionloss"])[0]
    assert nsc >= 2 + len(fl_dofs) and list(fl_dofs) == [1, 2, 9, 10, 17, 18]
    for r, (a, b) in enumerate(((6, 7), (14, 15))):                      # finger coupling rows
        assert np.array_equal(np.nonzero(J[r])[0], [a, b]) and np.allclose(J[r][[a, b]], [1, -1])
        assert abs(o.qpos[a] - o.qpos[b]) <= 2e-4                        # soft equality: the fingers stay within 0.2 mm
    moving = 0
    for r, d in enumerate(fl_dofs, start=2):
        fl = avm["dof_frictionloss"][d]
        assert np.array_equal(np.nonzero(J[r])[0], [d]) and abs(f[r]) <= fl + 1e-9
        if abs(o.qvel[d]) > 1e-2:
            moving += 1
            assert abs(f[r] + fl * np.sign(o.qvel[d])) <= 1e-6
    assert moving >= 3


def test_contact_parameters_follow_mujocos_mixing_rules(slot_model_path):
    """contact dimension = max of the two geoms' condim, friction = element-wise max of their (sliding, torsional, rolling)
    coefficients, expanded to (mu, mu, torsion, roll, roll) -- with the MJCF numbers: the slot rails (friction 0.05,
    task_slot_insertion.xml:7-8) against the condim-6 `collision`-class table still slide at the table's coefficient."""
    from av_aloha_b200 import model_io
    avm, names = model_io.load_avm(slot_model_path), model_io.load_names("slot_insertion", 3)["geom"]
    o = _settled_oracle(slot_model_path)
    C = o.contacts()
    assert len(C) >= 8
    seen = set()
    for c in C:
        g1, g2 = int(c[13]), int(c[14])
        f = np.maximum(avm["geom_friction"][g1], avm["geom_friction"][g2])
        assert int(c[15]) == max(int(avm["geom_condim"][g1]), int(avm["geom_condim"][g2]))
        assert np.abs(c[17:22] - [f[0], f[0], f[1], f[2], f[2]]).max() <= 1e-12
        seen.add((names[g1], names[g2]))
    assert ("table", "stick") in seen and ("table", "slot-1") in seen
    gs, gt = names.index("slot-1"), names.index("table")
    assert avm["geom_friction"][gs][0] == pytest.approx(0.05) and avm["geom_friction"][gt][0] >= 0.05
    assert avm["geom_condim"][gt] == 6
