"""Config 4 (BASELINE.json): HookPackage-2Arms, B = 8192, contact-rich grasp states -- throughput of env.step at the converged
solver setting (Newton, 3e-7) next to the fixed-sweep PGS settings of round 1, plus what each setting's solve is worth:
|qacc - qacc_ref|inf / max(1, |qacc_ref|inf) of one forward pass on the 64 fixture grasp states against the fp64 oracle's Newton
(1e-13) on the GPU's own contact list (mobile dofs; tests/test_solver_newton.py).

    python tools/config4_hook.py [B]
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from av_aloha_b200 import capi, model_io, workload
from oracle.oracle import OracleModel
from test_solver_newton import _mobile_dofs, _oracle, _rel, _states

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
task, arms = "hook_package", 2
path = model_io.model_path(task, arms)
model = capi.Model(path, 0)
avm = model_io.load_avm(path)
free = model_io.load_names(task, arms)["free_joint"]
rng = np.random.default_rng(4)
lo, hi = avm["reset_lo"], avm["reset_hi"]
fp = lo[None] + (hi - lo)[None] * rng.random((B, len(free), 3))
pkg = fp[:, free.index("package_joint")]
acts = np.tile(workload.HOME[:14], (B, 1))
for arm, sgn in ((0, 1.0), (1, -1.0)):
    n = 6
    w0, p0, site0 = avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm]
    Rt = workload._roty(sgn * 1.0) @ site0[:3, :3]
    off = workload._roty(sgn * 1.0) @ np.array([sgn * workload.PAD_FWD, 0.0, -workload.PAD_DOWN])
    q, err = workload.solve_ik(np.tile(workload.HOME[7 * arm:7 * arm + 6], (B, 1)), pkg + np.array([-sgn * 0.03, 0.0, 0.03]) - off,
                               np.broadcast_to(Rt, (B, 3, 3)), w0, p0, site0, avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1])
    acts[:, 7 * arm:7 * arm + 6] = q
acts += rng.normal(0, 0.01, acts.shape)
acts[:, [6, 13]] = 0.2
acts = torch.as_tensor(acts.astype(np.float32), device="cuda")
st = _states(task)
om = OracleModel(path)
mob = _mobile_dofs(path)
settings = [("newton", 0)] + [("pgs", k) for k in (4, 8, 16, 32, 64)]
print(f"HookPackage-2Arms B={B}: both grippers closing on the package, 25 settle steps, then 10 timed env.steps per setting")
for solver, iters in settings:
    batch = capi.Batch(model, B, seed=4)
    batch.set_solver(solver)
    if solver == "pgs":
        batch.set_options(solver_iters=iters)
        batch.set_warmstart(2)
    batch.reset(free_pos=fp)
    for _ in range(25):
        batch.step(acts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        batch.step(acts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ncon = batch.get(capi.NCON).float().mean().item()
    bad = int((batch.get(capi.STATUS) & 1).sum().item())
    nst = batch.get(capi.SOLVER_STAT)[:, 0].mean().item() / 20
    batch.close()
    # accuracy of this setting's solve on the fixture states
    b2 = capi.Batch(model, len(st["qpos"]))
    b2.set_solver(solver)
    if solver == "pgs":
        b2.set_options(solver_iters=iters)
    for k, f in (("qpos", capi.QPOS), ("qvel", capi.QVEL), ("ctrl", capi.CTRL), ("warm", capi.WARMSTART)):
        b2.set(f, st[k])
    b2.forward()
    qacc, nc, con = b2.get(capi.QACC).cpu().numpy(), b2.get(capi.NCON).cpu().numpy(), b2.get(capi.CONTACTS).cpu().numpy()
    b2.close()
    errs = []
    for e in range(len(st["qpos"])):
        c = con[e][: nc[e]]
        o = _oracle(om, st, e, contacts=np.concatenate([c[:, 0:7], c[:, 7:9]], axis=1))
        errs.append(_rel(qacc[e][mob], o.qacc[mob]))
    errs = np.array(errs)
    tag = "Newton 3e-7" if solver == "newton" else f"PGS {iters:2d} sweeps (cold forward pass)"
    print(f"  {tag:32s}: {ms:7.2f} ms/step = {B / ms * 1e3:8.0f} env-steps/s  ncon {ncon:.1f} blown-up {bad}"
          + (f" newton iters/substep {nst:.2f}" if solver == "newton" else "")
          + f" | rel |dqacc| median {np.median(errs):.1e} p90 {np.quantile(errs, .9):.1e} max {errs.max():.1e}")
