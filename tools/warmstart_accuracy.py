"""Accuracy of the PGS solve along the bench workload for the two warm-start schemes, measured with the fp64 oracle (CPU).

For every setting, 6 environments of the SlotInsertion scripted-policy workload are stepped for 200 env.steps (reach,
grasp, lift, carry) and compared, every 20 steps, with the same environments solved with 400 sweeps per substep
(|qpos - qpos_400|_inf, median over environments).  Contact dynamics are chaotic, so the curves diverge eventually; what
matters is the ordering of the settings.

    python tools/warmstart_accuracy.py      # ~15 min on 8 cores
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from concurrent.futures import ThreadPoolExecutor
import numpy as np
from av_aloha_b200 import model_io, workload
from oracle.oracle import OracleEnv, OracleModel

path = model_io.model_path("slot_insertion", 3)
om = OracleModel(path)
N, T = 6, 200
obj = workload.sample_object_positions(N, 1234)
acts = workload.slot_insertion_script(300, obj, 1234).astype(np.float64)


def run(ws, iters):
    def one(e):
        o = OracleEnv(om)
        o.set_options(max_iter=iters, tol=0.0, warmstart=ws)
        o.reset(free_pos=obj[e])
        tr = []
        for t in range(T):
            o.step(acts[t, e])
            if t % 20 == 19:
                tr.append(o.qpos.copy())
        return np.stack(tr)
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        return np.stack(list(ex.map(one, range(N))))


ref = run(1, 400)
print("warm start            sweeps  median |dqpos|inf vs the 400-sweep run at t = 19, 39, ..., 199")
for ws, iters in ((1, 20), (1, 8), (2, 20), (2, 12), (2, 8), (2, 4)):
    err = np.abs(run(ws, iters) - ref).max(axis=2)
    name = "qacc map (MuJoCo)" if ws == 1 else "force cache"
    print(f"{name:20s}  {iters:5d}  " + " ".join(f"{v:.4f}" for v in np.median(err, axis=0)), flush=True)
