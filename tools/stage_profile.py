"""Per-stage SM-cycle breakdown of the step kernel on the steady-state bench workload
(needs the -DAVSIM_PROFILE build: tools/build_prof.sh).  python tools/stage_profile.py [B] [iters] [nsteps]"""
import ctypes as C
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ.setdefault("AVSIM_LIB", os.path.join(ROOT, "av_aloha_b200", "csrc", "libavsim_prof.so"))
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import steady
from av_aloha_b200 import capi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
NAMES = ["load", "kinematics", "inertia", "broadphase", "prim_narrow", "convex_narrow", "smooth", "rows_scalar",
         "rows_contact", "solve", "integrate", "outputs", "newton_init", "newton_grad", "newton_hess", "newton_chol", "newton_ls", "newton_wait"]
model, batch, acts, masks, mask_any, fp, t0 = steady.restore(B, iters)
batch.set_solver(os.environ.get('SOLVER', 'newton'))
lib = capi.load_library()
buf = (C.c_uint64 * 32)()
steady.step(batch, acts, masks, mask_any, fp, t0)
torch.cuda.synchronize()
lib.avsim_stage_cycles(buf, 32, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(nsteps):
    steady.step(batch, acts, masks, mask_any, fp, t0 + 1 + k)
e1.record()
torch.cuda.synchronize()
n = lib.avsim_stage_cycles(buf, 32, 1)
cyc = np.array(buf[:n], dtype=np.float64)
ncon = batch.get(capi.NCON).float()
print(f"B={B} iters={iters}: {e0.elapsed_time(e1) / nsteps:.1f} ms/step; ncon mean {ncon.mean().item():.1f} max {int(ncon.max().item())}; "
      f"reward mean {batch.get(capi.REWARD).float().mean().item():.2f}; status or {int(batch.get(capi.STATUS).max().item())}")
for k in range(n):
    print(f"   {NAMES[k]:14s} {cyc[k] / B / 20 / nsteps:10.0f} warp-resident cycles/env-substep  {100 * cyc[k] / cyc.sum():5.1f} %")
