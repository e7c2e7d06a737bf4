"""GPU parity on the other task models and on edge cases of the step path (oracle = fp64 CPU restatement)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0])


def _hold(nj):
    a = HOME[:nj].copy()
    a[6] = a[13] = 1.0
    return a


@pytest.mark.parametrize("task,arms", [("insert_peg", 2), ("sew_needle", 3), ("tube_transfer", 3), ("hook_package", 2)])
def test_env_step_parity_other_tasks(task, arms):
    import torch
    from av_aloha_b200 import capi, env, model_io
    from oracle.oracle import OracleEnv, OracleModel

    path = model_io.model_path(task, arms)
    model, om = capi.Model(path, 0), OracleModel(path)
    B = 3
    rng = np.random.default_rng(7)
    free = model_io.load_names(task, arms)["free_joint"]
    fp = np.stack([env.reference_reset_draws(task, free, rng) for _ in range(B)])
    b = capi.Batch(model, B, seed=1)
    b.set_solver("pgs"); b.set_options(solver_iters=60)
    b.reset(free_pos=fp)
    nj = model.njoints
    act = np.tile(_hold(nj), (B, 1)).astype(np.float32)
    act_dev = torch.as_tensor(act, device="cuda")
    oras = []
    for e in range(B):
        o = OracleEnv(om)
        o.set_solver("pgs"); o.set_options(max_iter=60, tol=0.0)
        o.reset(free_pos=fp[e])
        oras.append(o)
    for step in range(2):
        b.step(act_dev)
        qpos, rew, ncon = (b.get(f).cpu().numpy() for f in (capi.QPOS, capi.REWARD, capi.NCON))
        assert (b.get(capi.STATUS).cpu().numpy() == 0).all()
        for e in range(B):
            full = np.concatenate([act[e], HOME[nj:]]).astype(np.float64)
            r = oras[e].step(full)
            assert ncon[e] == oras[e].ncon and rew[e] == r, (task, step, e)
            assert np.abs(qpos[e] - oras[e].qpos).max() <= 1e-4, (task, step, e)
    b.close()


def test_zero_substeps_and_single_env_and_ragged_batch():
    """edge cases: nsubsteps = 0 leaves the state untouched, B = 1, and a batch size that is not a multiple of the
    block shape (the queue hands out partial groups)"""
    import torch
    from av_aloha_b200 import capi, model_io

    model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
    act1 = torch.as_tensor(_hold(21)[None].astype(np.float32), device="cuda")
    b = capi.Batch(model, 1, seed=2)
    q0 = b.get(capi.QPOS).clone()
    b.step(act1, nsubsteps=0)
    assert torch.equal(b.get(capi.QPOS), q0)
    b.step(act1)
    assert torch.isfinite(b.get(capi.QPOS)).all() and int(b.get(capi.STATUS).item()) == 0
    b.close()
    B = 45
    fp = np.tile(np.array([[[0.0, 0.12, 0.0], [0.0, -0.05, 0.0]]]), (B, 1, 1))
    b = capi.Batch(model, B, seed=2)
    b.reset(free_pos=fp)
    acts = torch.as_tensor(np.tile(_hold(21), (B, 1)).astype(np.float32), device="cuda")
    for _ in range(2):
        b.step(acts)
    q = b.get(capi.QPOS)
    assert float((q - q[0]).abs().max()) == 0.0           # identical inputs -> bit-identical rows, whatever warp ran them
    b.close()


def test_error_paths():
    from av_aloha_b200 import capi, model_io
    with pytest.raises(capi.AvsimError):
        capi.Model("/nonexistent/model.avm", 0)
    model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
    with pytest.raises(capi.AvsimError):
        capi.Batch(model, 0)
    b = capi.Batch(model, 2)
    import torch
    with pytest.raises(capi.AvsimError):
        b.set(capi.REWARD, torch.zeros(2, dtype=torch.int32))     # read-only field
    b.close()


def test_sew_needle_latch_and_reward_tables():
    """reward staging from contact classes: force contact sets by placing objects, compare with the oracle's rewards"""
    import torch
    from av_aloha_b200 import capi, model_io
    from oracle.oracle import OracleEnv, OracleModel

    path = model_io.model_path("sew_needle", 3)
    model, om = capi.Model(path, 0), OracleModel(path)
    rng = np.random.default_rng(3)
    B = 16
    b = capi.Batch(model, B, seed=5)
    b.reset()
    q = b.get(capi.QPOS).cpu().numpy()
    # scatter the needle / wall (free joints) over and above the table and perturb arm joints: a spread of contact sets
    fq = model.table("free_qadr")
    for k in range(model.nfree):
        q[:, fq[k]:fq[k] + 3] += rng.normal(0, 0.03, size=(B, 3)) * [1, 1, 0.3]
    q[:, :16] += rng.normal(0, 0.15, size=(B, 16))
    b.set(capi.QPOS, q.astype(np.float32))
    b.forward()
    q32 = b.get(capi.QPOS).cpu().numpy()
    act = torch.as_tensor(np.tile(_hold(21), (B, 1)).astype(np.float32), device="cuda")
    b.step(act, nsubsteps=0)      # position pass + reward only
    rew, latch, ncon = (b.get(f).cpu().numpy() for f in (capi.REWARD, capi.LATCH, capi.NCON))
    for e in range(B):
        o = OracleEnv(om)
        o.reset()
        o.qpos[:] = q32[e]
        r = o.step(_hold(21), nsub=0)
        assert ncon[e] == o.ncon and rew[e] == r and latch[e] == o.latch, e
    b.close()
