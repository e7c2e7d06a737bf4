"""Config 4 (BASELINE.json): HookPackage-2Arms, B = 8192, PGS-iteration sweep.

For solver_iters in {2,4,8,16,32,64} (+ 3 noslip sweeps fixed): env-steps/s of the batch, and the accuracy of one
forward pass -- max |qacc - qacc_ref| over 64 sampled contact-rich states against the fp64 oracle run to convergence
(3000 sweeps, tol 1e-14), relative to max(1, |qacc_ref|).  States: both grippers lowered onto the package (package on
the table, fingers closing on it), per-env joint noise, after 10 env.steps of settling.

    python tools/pgs_sweep.py [B] [nstates]
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from av_aloha_b200 import capi, model_io, workload
from oracle.oracle import OracleEnv, OracleModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 64
path = model_io.model_path("hook_package", 2)
model = capi.Model(path, 0)
avm = model_io.load_avm(path)
rng = np.random.default_rng(4)
free = model_io.load_names("hook_package", 2)["free_joint"]
lo, hi = avm["reset_lo"], avm["reset_hi"]
fp = lo[None] + (hi - lo)[None] * rng.random((B, len(free), 3))
pkg = fp[:, free.index("package_joint")]
# both hands reach for the package from their side (pads 3 cm above the table), then close
acts = np.tile(workload.HOME[:14], (B, 1))
for arm, sgn in ((0, 1.0), (1, -1.0)):
    n = 6
    w0, p0, site0 = avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm]
    Rt = workload._roty(sgn * 1.0) @ site0[:3, :3]
    off = workload._roty(sgn * 1.0) @ np.array([sgn * workload.PAD_FWD, 0.0, -workload.PAD_DOWN])
    tgt = pkg + np.array([-sgn * 0.03, 0.0, 0.03])
    q, err = workload.solve_ik(np.tile(workload.HOME[7 * arm:7 * arm + 6], (B, 1)), tgt - off, np.broadcast_to(Rt, (B, 3, 3)),
                               w0, p0, site0, avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1])
    acts[:, 7 * arm:7 * arm + 6] = q
acts += rng.normal(0, 0.01, acts.shape)
acts[:, [6, 13]] = 0.2
acts = torch.as_tensor(acts.astype(np.float32), device="cuda")
batch = capi.Batch(model, B, seed=4)
batch.set_options(solver_iters=32)
batch.reset(free_pos=fp)
for _ in range(25):
    batch.step(acts)
state = {f: batch.get(f).clone() for f in (capi.QPOS, capi.QVEL, capi.CTRL, capi.WARMSTART)}
print(f"B={B}: settled; ncon mean {batch.get(capi.NCON).float().mean().item():.1f} max {int(batch.get(capi.NCON).max().item())} "
      f"status or {int(batch.get(capi.STATUS).max().item())}")
idx = np.argsort(-batch.get(capi.NCON).cpu().numpy())[:NS]          # the most contact-rich states
om = OracleModel(path)
ref = []
q64, v64, c64, w64 = (state[f].cpu().numpy().astype(np.float64) for f in (capi.QPOS, capi.QVEL, capi.CTRL, capi.WARMSTART))
for e in idx:
    o = OracleEnv(om)
    o.set_options(max_iter=3000, tol=1e-14)
    o.reset(free_pos=fp[e])
    o.qpos[:], o.qvel[:], o.ctrl[:], o.qacc_warmstart[:] = q64[e], v64[e], c64[e], w64[e]
    o.forward()
    ref.append(o.qacc.copy())
ref = np.stack(ref)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
print("iters  env-steps/s   max rel |dqacc|   median rel |dqacc|")
for iters in (2, 4, 8, 16, 32, 64):
    batch.set_options(solver_iters=iters)
    for f, v in state.items():
        batch.set(f, v)
    batch.forward()
    qacc = batch.get(capi.QACC).cpu().numpy()[idx]
    rel = np.abs(qacc - ref).max(axis=1) / np.maximum(1.0, np.abs(ref).max(axis=1))
    for _ in range(2):
        batch.step(acts)
    e0.record()
    for _ in range(4):
        batch.step(acts)
    e1.record(); torch.cuda.synchronize()
    print(f"{iters:5d}  {B * 4 / e0.elapsed_time(e1) * 1e3:11.0f}   {rel.max():14.3e}   {np.median(rel):14.3e}", flush=True)
