"""GPU forward pass with the Newton solver on contact-rich states of the bench workload vs the fp64 oracle's Newton run
to convergence: |dqacc|inf / max(1, |qacc|inf) per environment (SURVEY.md 8c tolerance: 1e-2)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

from av_aloha_b200 import capi, model_io
from oracle.oracle import OracleEnv, OracleModel

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
z = np.load(os.path.join(ROOT, "tools", "_steady", "steady_B4096.npz"))
path = model_io.model_path("slot_insertion", 3)
model = capi.Model(path, 0)
idx = np.random.default_rng(5).choice(4096, N, replace=False)
b = capi.Batch(model, N)
b.set_solver("newton")
for k, f in (("qpos", capi.QPOS), ("qvel", capi.QVEL), ("ctrl", capi.CTRL), ("warm", capi.WARMSTART)):
    b.set(f, z[k][idx])
b.forward()
qacc = b.get(capi.QACC).cpu().numpy()
ncon = b.get(capi.NCON).cpu().numpy()
st = b.get(capi.SOLVER_STAT).cpu().numpy()
om = OracleModel(path)
errs = []
for j, e in enumerate(idx):
    o = OracleEnv(om)
    o.qpos[:] = z["qpos"][e]; o.qvel[:] = z["qvel"][e]; o.ctrl[:] = z["ctrl"][e]; o.qacc_warmstart[:] = z["warm"][e]
    o.set_options(max_iter=100, tol=1e-14, warmstart=1); o.set_solver("newton")
    o.forward()
    qr = o.qacc.copy()
    errs.append(np.abs(qacc[j] - qr).max() / max(1, np.abs(qr).max()))
    if errs[-1] > 1e-2:
        print("env", e, "ncon", o.ncon, ncon[j], "err %.3e" % errs[-1], "dof", np.abs(qacc[j] - qr).argmax(), "stat", st[j])
errs = np.array(errs)
print(f"N={N} rel err median {np.median(errs):.2e} p90 {np.quantile(errs,.9):.2e} p99 {np.quantile(errs,.99):.2e} max {errs.max():.2e} "
      f"frac>1e-2 {np.mean(errs>1e-2):.4f}; newton iters mean {st[:,0].mean():.2f} max {st[:,0].max():.0f}; grad max {st[:,1].max():.2e}")
