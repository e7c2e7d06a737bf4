"""Step the steady-state bench workload n times (for ncu captures / drift checks): python tools/run_steps.py B iters nsteps"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import steady
from av_aloha_b200 import capi

B, iters, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model, batch, acts, masks, mask_any, fp, t0 = steady.restore(B, iters)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
for k in range(n):
    ev[k][0].record()
    steady.step(batch, acts, masks, mask_any, fp, t0 + k)
    ev[k][1].record()
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
print("per-step ms:", " ".join(f"{x:.1f}" for x in ms))
print(f"done {n} steps, {sum(ms) / n:.1f} ms/step; ncon mean", batch.get(capi.NCON).float().mean().item())
