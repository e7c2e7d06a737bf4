"""Long-horizon check of the CUDA path at the bench's solver setting (8 sweeps + 3 noslip, force cache) against the fp64
oracle: 6 environments of the bench workload, 200 env.steps; |qpos_gpu - qpos_oracle|_inf every 20 steps against
(a) the oracle at the same setting (parity over a long horizon; contact dynamics are chaotic, so this grows) and
(b) the oracle with 400 sweeps per substep (how far the bench setting is from a converged solve).
    python tools/gpu_trajectory_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from concurrent.futures import ThreadPoolExecutor
import numpy as np, torch
from av_aloha_b200 import capi, model_io, workload
from oracle.oracle import OracleEnv, OracleModel

N, T = 6, 200
path = model_io.model_path("slot_insertion", 3)
obj = workload.sample_object_positions(N, 1234)
acts = workload.slot_insertion_script(300, obj, 1234)
model = capi.Model(path, 0)
b = capi.Batch(model, N, seed=1)
b.set_options(solver_iters=8); b.set_warmstart(2)
b.reset(free_pos=obj)
gpu = []
a_dev = torch.as_tensor(acts[:, :N], device="cuda")
for t in range(T):
    b.step(a_dev[t].contiguous())
    if t % 20 == 19:
        gpu.append(b.get(capi.QPOS).cpu().numpy().astype(np.float64))
gpu = np.stack(gpu, 1)
om = OracleModel(path)
def run(ws, iters):
    def one(e):
        o = OracleEnv(om); o.set_options(max_iter=iters, tol=0.0, warmstart=ws); o.reset(free_pos=obj[e])
        tr = []
        for t in range(T):
            o.step(acts[t, e].astype(np.float64))
            if t % 20 == 19: tr.append(o.qpos.copy())
        return np.stack(tr)
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        return np.stack(list(ex.map(one, range(N))))
same, conv = run(2, 8), run(1, 400)
print("t =                          " + " ".join(f"{t:6d}" for t in range(19, T, 20)))
print("vs oracle, same setting  max " + " ".join(f"{v:6.4f}" for v in np.abs(gpu - same).max(axis=(0, 2))))
print("vs oracle, same setting  med " + " ".join(f"{v:6.4f}" for v in np.median(np.abs(gpu - same).max(axis=2), axis=0)))
print("vs oracle, 400 sweeps    med " + " ".join(f"{v:6.4f}" for v in np.median(np.abs(gpu - conv).max(axis=2), axis=0)))
print("oracle same vs 400       med " + " ".join(f"{v:6.4f}" for v in np.median(np.abs(same - conv).max(axis=2), axis=0)))
print("rewards gpu", b.get(capi.REWARD).cpu().numpy())
