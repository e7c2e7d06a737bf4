"""Time block-shape / sync variants of the step kernel on the same mid-episode state:
   python tools/variant_bench.py [B] [iters] [t0] -- variants are (AVSIM_WARPS, AVSIM_SYNC) pairs."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from av_aloha_b200 import capi, model_io

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
t0 = int(sys.argv[3]) if len(sys.argv) > 3 else 150
variants = [tuple(int(x) for x in v.split(':')) for v in (sys.argv[4].split(',') if len(sys.argv) > 4 else ['1:0'])]
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
acts = torch.as_tensor(bench.script_actions(300, B, 1234), device="cuda")


def make(w, s):
    os.environ["AVSIM_WARPS"], os.environ["AVSIM_SYNC"] = str(w), str(s)
    b = capi.Batch(model, B, seed=1234)
    b.set_options(solver_iters=iters)
    return b


b0 = make(1, 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(t0):
    b0.step(acts[t])
e1.record(); torch.cuda.synchronize()
print(f"advance to t={t0}: {e0.elapsed_time(e1) / t0:.1f} ms/step avg (warps=1 sync=0)")
state = {f: b0.get(f).clone() for f in (capi.QPOS, capi.QVEL, capi.CTRL, capi.WARMSTART)}
ref = None
for w, s in variants:
    b = make(w, s)
    for f, v in state.items():
        b.set(f, v)
    ts = []
    for k in range(4):
        e0.record(); b.step(acts[t0 + k]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    q = b.get(capi.QPOS)
    if ref is None:
        ref = q.clone()
    print(f"warps={w} sync={s}: {sum(ts[1:]) / 3:7.1f} ms/step  -> {B / (sum(ts[1:]) / 3) * 1e3:8.0f} env-steps/s   "
          f"max|dqpos| vs first variant {float((q - ref).abs().max()):.2e}  ncon {b.get(capi.NCON).float().mean().item():.1f}", flush=True)
    b.close()
