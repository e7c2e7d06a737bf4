"""BASELINE.json's full size (configs[1]: SlotInsertion-3Arms, 4096 environments, the bench's scripted grasp / lift / insert
workload) through size-independent properties -- the oracle cannot run 4096 x 300 steps, but these must hold at any size:

  * determinism: two batches given the same reset draws and actions agree bit for bit after every step, although the split
    pipeline runs three environment groups on three streams with cost-sorted queues whose order changes from step to step;
  * composition independence: environment e of the 4096-batch steps exactly like the same environment in a batch of 64;
  * invariants of the state: nothing blows up, free-body quaternions stay unit, the objects do not sink through the table, the
    joints stay inside their ranges (plus the solver's slack), rewards stay in 0..max, contact counts within the 64 slots;
  * the solve converges everywhere: no Newton solve at the iteration cap, scaled gradient at the stop rule for all but a
    handful of the 4096 x 20 x steps solves.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_batch_determinism_composition_and_invariants(monkeypatch):
    import torch
    from av_aloha_b200 import capi, model_io, workload

    task, B, T, K = "slot_insertion", 4096, 120, 64
    path = model_io.model_path(task, 3)
    model = capi.Model(path, 0)
    avm = model_io.load_avm(path)
    obj = workload.sample_object_positions(B, 11)
    acts = workload.slot_insertion_script(300, obj, 11)                   # [300, B, 21]
    phase = (np.arange(B) * 40) // B                                       # episodes staggered over 40 steps: reach, pinch, grasp
    idx = lambda t: np.minimum(t + phase, 299)
    sel = np.linspace(0, B - 1, K).astype(int)

    monkeypatch.setenv("AVSIM_SPLIT", "1")      # the 64-environment batch would otherwise run the fused kernel (same math, other rounding)

    def run(envs):
        b = capi.Batch(model, len(envs), seed=3)
        assert b.launch_shape["split"] == 1
        b.reset(free_pos=obj[envs])
        out = []
        for t in range(T):
            a = acts[idx(t)[envs], envs]
            b.step(torch.as_tensor(np.ascontiguousarray(a), device="cuda"))
            if t % 40 == 39:
                out.append({k: b.get(f).cpu().numpy().copy() for k, f in (("qpos", capi.QPOS), ("qvel", capi.QVEL), ("reward", capi.REWARD),
                                                                          ("ncon", capi.NCON), ("status", capi.STATUS), ("stat", capi.SOLVER_STAT))})
        b.close()
        return out

    allenv = np.arange(B)
    a, b2, small = run(allenv), run(allenv), run(sel)
    for x, y, s in zip(a, b2, small):
        for k in ("qpos", "qvel", "reward", "ncon"):
            assert np.array_equal(x[k], y[k]), k                           # determinism
            assert np.array_equal(x[k][sel], s[k]), k                      # composition independence
    last = a[-1]
    assert (last["status"] == 0).all()
    assert last["ncon"].max() <= 64 and last["ncon"].mean() > 5            # the workload is in contact
    assert last["reward"].min() >= 0 and last["reward"].max() <= 4
    q = last["qpos"].astype(np.float64)
    assert np.isfinite(q).all() and np.isfinite(last["qvel"]).all()
    for adr in (int(a_) for a_ in avm["free_qadr"]):
        assert np.abs(np.linalg.norm(q[:, adr + 3: adr + 7], axis=1) - 1.0).max() <= 1e-5      # unit quaternions
        assert q[:, adr + 2].min() >= -2e-3                                 # not through the table (z of the body origin)
    lo, hi = avm["jnt_range"][:, 0], avm["jnt_range"][:, 1]
    hinge = np.nonzero(avm["jnt_limited"] != 0)[0]
    qa = q[:, avm["jnt_qposadr"][hinge]]
    assert (qa >= lo[hinge] - 0.02).all() and (qa <= hi[hinge] + 0.02).all()                    # joint limits (soft: solref slack)
    for x in a:
        st = x["stat"]                                                      # [B, 4]: iterations, worst scaled gradient, max iterations, solves at the cap
        assert st[:, 3].sum() <= 2 and st[:, 2].max() <= 30
        assert np.mean(st[:, 1] > 1e-5) <= 1e-3
