"""Run the bench workload and stop at the first environment whose status bit 0 (numerical blow-up) or 4 (Hessian not PD) is set;
saves the state that environment had BEFORE the offending step to gpurun_out/blowup.npz (replayed on the CPU afterwards)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from av_aloha_b200 import capi, model_io

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 700
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
batch = capi.Batch(model, B, seed=1234)
obj, acts_np, masks_np, phase = bench.make_workload(B, 1234)
acts = torch.as_tensor(acts_np, device="cuda")
masks = torch.as_tensor(masks_np, device="cuda")
fp = torch.as_tensor(obj.astype(np.float32), device="cuda")
mask_any = masks_np.any(axis=1)
batch.reset(free_pos=fp)
found = []
for k in range(nsteps):
    t = k % bench.EPISODE_LEN
    if mask_any[t]:
        batch.reset(mask=masks[t], free_pos=fp)
    prev = {n: batch.get(f).clone() for n, f in (("qpos", capi.QPOS), ("qvel", capi.QVEL), ("ctrl", capi.CTRL), ("warm", capi.WARMSTART))}
    batch.step(acts[t])
    st = batch.get(capi.STATUS)
    bad = torch.nonzero((st & 17) != 0).flatten().cpu().numpy()
    new = [int(e) for e in bad if int(e) not in [f[0] for f in found]]
    for e in new:
        found.append((e, k))
        print("step", k, "env", e, "status", int(st[e].item()), "solver stat", batch.get(capi.SOLVER_STAT)[e].cpu().numpy(), flush=True)
        np.savez(os.path.join(ROOT, "gpurun_out", f"blowup_{len(found)}.npz"), env=e, step=k, action=acts_np[t, e],
                 **{n: v[e].cpu().numpy() for n, v in prev.items()})
    if len(found) >= 4:
        break
print("done", found)
