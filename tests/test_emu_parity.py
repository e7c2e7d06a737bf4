"""CPU-only check of the step kernel's SOURCE: csrc/*.cuh compiled for the host by the warp-emulation harness
(tests/emu, 32 cooperative fibers per warp) against the fp64 oracle.  Same tolerances as the GPU parity tests
(tests/test_gpu_parity.py); the harness is test infrastructure, not a product path."""
import numpy as np
import pytest

HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0])


@pytest.fixture(scope="module")
def ctx(slot_model_path):
    from oracle.oracle import OracleEnv, OracleModel
    from tests.emu.emu import EmuBatch
    return EmuBatch, OracleModel(slot_model_path), OracleEnv, slot_model_path


def _fp(B, rng):
    return np.stack([np.array([[rng.uniform(-0.05, 0.05), rng.uniform(0.1, 0.15), 0.0],
                               [rng.uniform(-0.08, 0.08), rng.uniform(-0.1, 0.0), 0.0]]) for _ in range(B)])


def test_forward_random_state_converged(ctx):
    """random joint offsets + velocities, objects resting on the table: contacts, equality, friction-loss rows all live"""
    EmuBatch, om, OracleEnv, path = ctx
    B = 4
    rng = np.random.default_rng(0)
    fp = _fp(B, rng)
    eb = EmuBatch(path, B)
    eb.set_solver("pgs"); eb.set_options(150)
    eb.reset(fp)
    eb.qpos[:, :23] += rng.normal(0, 0.05, size=(B, 23)).astype(np.float32)
    eb.qvel[:] = rng.normal(0, 0.3, size=(B, eb.nv))
    eb.forward()
    for e in range(B):
        o = OracleEnv(om)
        o.reset(free_pos=fp[e])
        o.qpos[:] = eb.qpos[e]
        o.qvel[:] = eb.qvel[e]
        o.forward()
        assert o.ncon == eb.ncon[e]
        scale = max(1.0, np.abs(o.qacc).max())
        tol = 1e-2 if o.ncon else 1e-4
        assert np.abs(eb.qfrc_bias[e] - o.qfrc_bias).max() <= 1e-4 * max(1.0, np.abs(o.qfrc_bias).max())
        assert np.abs(eb.qacc[e] - o.qacc).max() <= tol * scale


def test_one_env_step_resting_contacts(ctx):
    EmuBatch, om, OracleEnv, path = ctx
    fp = np.array([[[0.01, 0.12, 0.0], [0.02, -0.05, 0.0]]])
    eb = EmuBatch(path, 1)
    eb.set_solver("pgs"); eb.set_options(50)
    eb.reset(fp)
    act = HOME.copy()
    act[6] = act[13] = 1.0
    o = OracleEnv(om)
    o.reset(free_pos=fp[0])
    eb.step(act[None].astype(np.float32), 20)
    r = o.step(act)
    assert eb.ncon[0] == o.ncon and eb.reward[0] == r and eb.status[0] == 0
    assert np.abs(eb.qpos[0] - o.qpos).max() <= 1e-4


def test_mixed_primitive_and_convex_contacts_keep_oracle_order(ctx):
    """stick wedged under the left fingers (mesh-hull contacts, MPR + multiccd) while the slot rests on the table
    (box-box contacts): same contact list in the same order as the oracle, and one env.step at a fixed, unconverged
    sweep count agrees -- which it only can if both walk the contacts in the same order"""
    EmuBatch, om, OracleEnv, path = ctx
    fp = np.array([[[0.0, 0.12, -0.002], [0.06, -0.011, 0.133]]])
    eb = EmuBatch(path, 1)
    eb.set_solver("pgs"); eb.set_options(30)
    eb.reset(fp)
    o = OracleEnv(om)
    o.set_solver("pgs"); o.set_options(max_iter=30, tol=0.0)
    o.reset(free_pos=fp[0])
    o.forward()
    assert o.ncon == eb.ncon[0] and o.ncon >= 6
    oc = o.contacts()
    ec = eb.contacts[0].reshape(-1, 16)[: o.ncon]
    assert np.array_equal(ec[:, 7:9].astype(int), oc[:, 13:15].astype(int))          # geom pairs, in order
    mesh = set(np.nonzero(np.array([7 == t for t in _geom_types(path)]))[0])
    assert any(int(g) in mesh for g in oc[:, 13]) and any(int(g) not in mesh and int(h) not in mesh for g, h in oc[:, 13:15])
    assert np.abs(ec[:, 0] - oc[:, 0]).max() <= 1e-5 and np.abs(ec[:, 1:4] - oc[:, 1:4]).max() <= 1e-4
    act = HOME.copy()
    act[6] = act[13] = 1.0
    eb.step(act[None].astype(np.float32), 20)
    r = o.step(act)
    assert eb.ncon[0] == o.ncon and eb.reward[0] == r
    assert np.abs(eb.qpos[0] - o.qpos).max() <= 2e-4


def _geom_types(path):
    from av_aloha_b200 import model_io
    return model_io.load_avm(path)["geom_type"]


def test_force_cache_warm_start_matches_oracle(ctx):
    """warm start mode 2 (each constraint starts from the force it carried in the previous solve, matched by identity):
    kernel source and oracle agree at a low, unconverged sweep count over two env.steps with mixed contacts"""
    EmuBatch, om, OracleEnv, path = ctx
    fp = np.array([[[0.0, 0.12, -0.002], [0.06, -0.011, 0.133]]])
    eb = EmuBatch(path, 1)
    eb.set_solver("pgs"); eb.set_options(8)
    eb.set_warmstart(2)
    eb.reset(fp)
    o = OracleEnv(om)
    o.set_solver("pgs"); o.set_options(max_iter=8, tol=0.0, warmstart=2)
    o.reset(free_pos=fp[0])
    act = HOME.copy()
    act[6] = act[13] = 0.3
    for _ in range(2):
        eb.step(act[None].astype(np.float32), 20)
        r = o.step(act)
    assert eb.ncon[0] == o.ncon and eb.reward[0] == r and eb.status[0] == 0
    assert np.abs(eb.qpos[0] - o.qpos).max() <= 2e-4


def test_action_to_ctrl_and_agent_pos_follow_the_reference_maps(ctx):
    """reference env.py:156-161 (gripper ctrlrange affine maps), 169-178 (agent_pos = L[6 joints + normalised gripper], R, M) and
    204-215 (ctrl write): a zero-substep env.step through the CUDA source shows exactly these maps.  The observed finger is
    left_left_finger for the left arm but right_RIGHT_finger for the right arm (constants.py:29-46), while both gripper
    actuators drive the *_left_finger joint (joint_position_actuators.xml:9,17)."""
    EmuBatch, om, OracleEnv, path = ctx
    lo, hi = 0.002, 0.037                                   # aloha_sim.xml:95 ctrlrange of the gripper actuators
    rng = np.random.default_rng(9)
    eb = EmuBatch(path, 1)
    eb.reset(np.array([[[0.0, 0.12, 0.0], [0.02, -0.05, 0.0]]]))
    a = rng.uniform(-0.5, 0.5, 21).astype(np.float32)
    a[6], a[13] = 0.3, 0.8                                  # 1 = open, 0 = closed
    eb.qpos[0, 6], eb.qpos[0, 15] = 0.010, 0.030            # the observed finger joints, somewhere inside their range
    eb.step(a[None], 0)
    want_ctrl = a.astype(np.float64).copy()
    want_ctrl[6], want_ctrl[13] = 0.3 * (hi - lo) + lo, np.float32(0.8) * (hi - lo) + lo
    assert np.abs(eb.ctrl[0] - want_ctrl).max() <= 1e-7
    q = eb.qpos[0].astype(np.float64)
    want_obs = np.concatenate([q[0:6], [(q[6] - lo) / (hi - lo)], q[8:14], [(q[15] - lo) / (hi - lo)], q[16:23]])
    assert np.abs(eb.agent_pos[0] - want_obs).max() <= 1e-6
    assert abs(eb.agent_pos[0, 6] - (0.010 - lo) / (hi - lo)) <= 1e-6 and abs(eb.agent_pos[0, 13] - 0.8) <= 1e-6


def test_reset_state_follows_the_reference_reset(ctx):
    """reference env.py:228-244 + constants.py:26-28: arm joints at the home poses, all four fingers at unnorm(1) = 0.037 (the
    0.02239 of the pose lists is overwritten by the gripper-joint bind that follows), ctrl = home with the gripper entries at
    0.037, velocities zero; objects where the host put them with identity orientation (env.py:536-537)."""
    EmuBatch, om, OracleEnv, path = ctx
    fp = np.array([[[0.03, 0.12, 0.0], [-0.02, -0.05, 0.0]]])
    eb = EmuBatch(path, 1)
    eb.qvel[0, :] = 1.0
    eb.reset(fp)
    arm, mid = [0, -0.082, 1.06, 0, -0.953, 0], [0, -0.8, 0.8, 0, 0.5, 0, 0]
    want_q = np.array(arm + [0.037, 0.037] + arm + [0.037, 0.037] + mid)
    assert np.abs(eb.qpos[0, :23] - want_q).max() <= 1e-7
    assert np.abs(eb.ctrl[0] - np.array(arm + [0.037] + arm + [0.037] + mid)).max() <= 1e-7
    assert np.abs(eb.qvel[0]).max() == 0.0
    assert np.abs(eb.qpos[0, 23:26] - fp[0, 0]).max() <= 1e-7 and np.abs(eb.qpos[0, 26:30] - [1, 0, 0, 0]).max() == 0.0
    assert np.abs(eb.qpos[0, 30:33] - fp[0, 1]).max() <= 1e-7 and np.abs(eb.qpos[0, 33:37] - [1, 0, 0, 0]).max() == 0.0
    assert np.abs(eb.agent_pos[0, [6, 13]] - 1.0).max() <= 1e-6 and eb.reward[0] == 0


def test_order_kernel_chunked_sort_is_a_heaviest_first_permutation(slot_model_path):
    """The queue-order kernel (source compiled for the host): one block sorts up to `chunk` environments by last step's cost,
    larger batches are sorted chunk by chunk and interleaved.  Checked: a permutation for every size, exactly sorted (ties in
    environment order) within one chunk, heaviest-first to within the interleaving for several chunks, ragged last chunk."""
    import ctypes as C
    from tests.emu import emu
    eb = emu.EmuBatch(slot_model_path, 1)
    L = emu.lib()
    L.emu_order.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int, C.c_int, C.POINTER(C.c_int)]
    rng = np.random.default_rng(0)
    for n, chunk in ((1, 64), (37, 64), (64, 64), (65, 64), (200, 64), (1000, 256), (513, 256)):
        cyc = rng.integers(0, 50, n).astype(np.int64) * 1000            # many ties
        order = np.full(n, -1, np.int32)
        L.emu_order(eb.ptr, cyc.ctypes.data_as(C.POINTER(C.c_longlong)), n, chunk, order.ctypes.data_as(C.POINTER(C.c_int)))
        assert sorted(order.tolist()) == list(range(n)), (n, chunk)     # every environment exactly once
        nch = (n + chunk - 1) // chunk
        if nch == 1:
            want = np.lexsort((np.arange(n), -cyc))                      # descending cost, ties in ascending environment order
            assert np.array_equal(order, want), (n, chunk)
        else:
            for c in range(nch):                                         # each chunk's environments appear in sorted order ...
                mine = [e for e in order if c * chunk <= e < (c + 1) * chunk]
                lo = c * chunk
                m = min(chunk, n - lo)
                want = lo + np.lexsort((np.arange(m), -cyc[lo:lo + m]))
                assert mine == want.tolist(), (n, chunk, c)
            first = order[:nch]                                          # ... and the queue starts with every chunk's heaviest
            assert sorted(first // chunk) == list(range(nch))
