"""Steady-state snapshot of the staggered bench workload, so profiling tools skip the 300-step pre-roll.

make : python tools/steady.py make [B] [iters]  -> gpurun_out/steady_B{B}.npz (copy it to tools/_steady/ to reuse it)
use  : batch, acts, masks, fp, t0 = steady.restore(model, B, iters)
"""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
from av_aloha_b200 import capi, model_io

FIELDS = {"qpos": capi.QPOS, "qvel": capi.QVEL, "ctrl": capi.CTRL, "warm": capi.WARMSTART, "latch": capi.LATCH,
          "fc_key": capi.FC_KEY, "fc_n": capi.FC_N, "fc_val": capi.FC_VAL}   # incl. the solver's force cache: without it the
# first solve after a restore starts from zero forces and the grasped objects slip (a lighter, unrepresentative state)


def path(B):
    return os.path.join(ROOT, "tools", "_steady", f"steady_B{B}.npz")


def setup(B, iters, seed=1234):
    model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
    batch = capi.Batch(model, B, seed=seed)
    batch.set_options(solver_iters=iters)
    batch.set_warmstart(int(os.environ.get("WARM", "2")))
    obj, acts_np, masks_np, phase = bench.make_workload(B, seed)
    acts = torch.as_tensor(acts_np, device="cuda")
    masks = torch.as_tensor(masks_np, device="cuda")
    fp = torch.as_tensor(obj.astype(np.float32), device="cuda")
    return model, batch, acts, masks, masks_np.any(axis=1), fp


def step(batch, acts, masks, mask_any, fp, t):
    t %= bench.EPISODE_LEN
    if mask_any[t]:
        batch.reset(mask=masks[t], free_pos=fp)
    batch.step(acts[t])


def restore(B, iters, seed=1234):
    model, batch, acts, masks, mask_any, fp = setup(B, iters, seed)
    p = path(B)
    if os.path.exists(p):
        z = np.load(p)
        batch.reset(free_pos=fp)
        for k, f in FIELDS.items():
            if k in z.files:
                batch.set(f, z[k])
        t0 = int(z["t"])
    else:
        batch.reset(free_pos=fp)
        for t in range(bench.EPISODE_LEN):
            step(batch, acts, masks, mask_any, fp, t)
        t0 = bench.EPISODE_LEN
    torch.cuda.synchronize()
    return model, batch, acts, masks, mask_any, fp, t0


if __name__ == "__main__":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    model, batch, acts, masks, mask_any, fp = setup(B, iters)
    batch.reset(free_pos=fp)
    for t in range(bench.EPISODE_LEN):
        step(batch, acts, masks, mask_any, fp, t)
    out = {k: batch.get(f).cpu().numpy() for k, f in FIELDS.items()}
    out["t"] = np.array(bench.EPISODE_LEN)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"steady_B{B}.npz"), **out)
    print("saved; ncon mean", batch.get(capi.NCON).float().mean().item(), "reward mean", batch.get(capi.REWARD).float().mean().item())
