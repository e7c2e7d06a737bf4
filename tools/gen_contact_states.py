"""Contact-rich states for the solver / collision parity tests -> tests/golden/contact_states.npz (run on a GPU box; the
states are INPUTS -- qpos, qvel, ctrl, previous qacc as float32 -- every expected value in the tests comes from the oracle).

  slot_insertion (3 arms): 256 states of the bench workload's steady state (staggered 300-step episodes, scripted
      grasp / lift / insert policy): the 192 most contact-rich of 4096 environments + 64 random ones;
  hook_package (2 arms, BASELINE config 4), sew_needle (3 arms): 64 grasp states each: both grippers lowered onto the first
      task object with per-environment joint noise and closing fingers, after 25 env.steps of settling (tools/pgs_sweep.py's recipe).

    python tools/gen_contact_states.py          # writes gpurun_out/contact_states.npz; copy it to tests/golden/
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from av_aloha_b200 import capi, model_io, workload

FIELDS = (("qpos", capi.QPOS), ("qvel", capi.QVEL), ("ctrl", capi.CTRL), ("warm", capi.WARMSTART))
out = {}


def grab(batch, idx, task):
    for n, f in FIELDS:
        out[f"{task}_{n}"] = batch.get(f).cpu().numpy()[idx].astype(np.float32)
    ncon = batch.get(capi.NCON).cpu().numpy()[idx]
    print(task, "states", len(idx), "ncon mean %.1f min %d max %d" % (ncon.mean(), ncon.min(), ncon.max()), "status or",
          int(batch.get(capi.STATUS).max().item()))


# ---- slot insertion: the bench workload after one full staggered episode
B = 4096
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
batch = capi.Batch(model, B, seed=1234)
obj, acts_np, masks_np, phase = bench.make_workload(B, 1234)
acts = torch.as_tensor(acts_np, device="cuda")
masks = torch.as_tensor(masks_np, device="cuda")
fp = torch.as_tensor(obj.astype(np.float32), device="cuda")
mask_any = masks_np.any(axis=1)
batch.reset(free_pos=fp)
for t in range(bench.EPISODE_LEN + 37):
    tt = t % bench.EPISODE_LEN
    if mask_any[tt]:
        batch.reset(mask=masks[tt], free_pos=fp)
    batch.step(acts[tt])
ncon = batch.get(capi.NCON).cpu().numpy()
rng = np.random.default_rng(7)
rich = np.argsort(-ncon, kind="stable")[:192]
rest = np.setdiff1d(np.arange(B), rich)
idx = np.concatenate([rich, rng.choice(rest, 64, replace=False)])
grab(batch, idx, "slot_insertion")
batch.close()

# ---- grasp states of two other tasks
for task, arms in (("hook_package", 2), ("sew_needle", 3)):
    B = 1024
    path = model_io.model_path(task, arms)
    model = capi.Model(path, 0)
    avm = model_io.load_avm(path)
    free = model_io.load_names(task, arms)["free_joint"]
    lo, hi = avm["reset_lo"], avm["reset_hi"]
    rng = np.random.default_rng(4)
    fpos = lo[None] + (hi - lo)[None] * rng.random((B, len(free), 3))
    target = fpos[:, free.index("package_joint" if task == "hook_package" else "needle_joint")]
    nj = model.njoints
    a = np.tile(workload.HOME[:nj], (B, 1))
    for arm, sgn in ((0, 1.0), (1, -1.0)):
        n = 6
        w0, p0, site0 = avm["ik_w0"][arm, :n], avm["ik_p0"][arm, :n], avm["ik_site0"][arm]
        Rt = workload._roty(sgn * 1.0) @ site0[:3, :3]
        off = workload._roty(sgn * 1.0) @ np.array([sgn * workload.PAD_FWD, 0.0, -workload.PAD_DOWN])
        tgt = target + np.array([-sgn * 0.03, 0.0, 0.03])
        q, err = workload.solve_ik(np.tile(workload.HOME[7 * arm:7 * arm + 6], (B, 1)), tgt - off, np.broadcast_to(Rt, (B, 3, 3)),
                                   w0, p0, site0, avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1])
        a[:, 7 * arm:7 * arm + 6] = q
    a += rng.normal(0, 0.01, a.shape)
    a[:, [6, 13]] = 0.2
    a = torch.as_tensor(a.astype(np.float32), device="cuda")
    batch = capi.Batch(model, B, seed=4)
    batch.reset(free_pos=fpos)
    for _ in range(25):
        batch.step(a)
    ncon = batch.get(capi.NCON).cpu().numpy()
    idx = np.argsort(-ncon, kind="stable")[:64]
    grab(batch, idx, task)
    batch.close()

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "contact_states.npz"), **out)
print("wrote gpurun_out/contact_states.npz")
