"""Per-stage SM-cycle breakdown of the step kernel (needs the -DAVSIM_PROFILE build: tools/build_prof.sh)."""
import ctypes as C
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ.setdefault("AVSIM_LIB", os.path.join(ROOT, "av_aloha_b200", "csrc", "libavsim_prof.so"))
import numpy as np
import torch

import bench
from av_aloha_b200 import capi, model_io

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 60, 110, 150, 250]
NAMES = ["load", "kinematics", "inertia", "broadphase", "prim_narrow", "convex_narrow", "smooth", "rows_scalar",
         "rows_contact", "solve", "integrate", "outputs"]
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
batch = capi.Batch(model, B, seed=1234)
batch.set_options(solver_iters=iters)
acts = torch.as_tensor(bench.script_actions(300, B, 1234), device="cuda")
lib = capi.load_library()
buf = (C.c_uint64 * 16)()
t = 0
for target in t_list:
    while t < target:
        batch.step(acts[t]); t += 1
    lib.avsim_stage_cycles(buf, 16, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); batch.step(acts[t]); e1.record(); t += 1
    torch.cuda.synchronize()
    n = lib.avsim_stage_cycles(buf, 16, 1)
    cyc = np.array(buf[:n], dtype=np.float64)
    ncon = batch.get(capi.NCON).float()
    print(f"t={t-1} B={B} iters={iters}: {e0.elapsed_time(e1):.1f} ms; ncon mean {ncon.mean().item():.1f} max {int(ncon.max().item())}; "
          f"reward max {int(batch.get(capi.REWARD).max().item())}; status or {int(batch.get(capi.STATUS).max().item())}")
    for k in range(n):
        print(f"   {NAMES[k]:14s} {cyc[k] / B / 20:10.0f} cycles/env-substep  {100 * cyc[k] / cyc.sum():5.1f} %")
