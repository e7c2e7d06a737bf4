"""Step the steady-state bench workload n times (for ncu captures): python tools/run_steps.py B iters nsteps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import steady
from av_aloha_b200 import capi

B, iters, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model, batch, acts, masks, mask_any, fp, t0 = steady.restore(B, iters)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(n):
    steady.step(batch, acts, masks, mask_any, fp, t0 + k)
e1.record()
torch.cuda.synchronize()
print(f"done {n} steps, {e0.elapsed_time(e1) / n:.1f} ms/step; ncon mean", batch.get(capi.NCON).float().mean().item())
