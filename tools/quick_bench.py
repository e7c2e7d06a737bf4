"""First-light timing: B envs holding the home pose (objects resting on the table)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from av_aloha_b200 import capi, model_io

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
batch = capi.Batch(model, B, seed=1234)
batch.set_options(solver_iters=iters)
HOME = np.array([0, -0.082, 1.06, 0, -0.953, 0, 0.02239] * 2 + [0, -0.8, 0.8, 0, 0.5, 0, 0], np.float32)
act = HOME.copy(); act[6] = 1; act[13] = 1
act = torch.as_tensor(np.tile(act, (B, 1)), device="cuda")
for _ in range(3):
    batch.step(act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    batch.step(act)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(f"B={B} iters={iters}: {ms:.2f} ms/step -> {B / ms * 1e3:.0f} env-steps/s; ncon mean {batch.get(capi.NCON).float().mean().item():.1f} "
      f"status {batch.get(capi.STATUS).max().item()} reward max {batch.get(capi.REWARD).max().item()}")
