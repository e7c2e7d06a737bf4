"""Step the scripted SlotInsertion batch n times (for ncu captures): python tools/run_steps.py B iters nsteps [t0]."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from av_aloha_b200 import capi, model_io

B, iters, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
batch = capi.Batch(model, B, seed=1234)
batch.set_options(solver_iters=iters)
acts = torch.as_tensor(bench.script_actions(300, B, 1234), device="cuda")
for t in range(n):
    batch.step(acts[t % 300])
torch.cuda.synchronize()
print("done", n, "steps; ncon mean", batch.get(capi.NCON).float().mean().item())
