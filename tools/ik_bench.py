"""Throughput of the IK kernels (not yet measured in round 1 -- run this first thing in round 2):

    python tools/ik_bench.py [n_problems=65536]

DiffIK (sim parameters, 10 iterations, 7-dof camera arm) and GradIK (sim parameters, <= 50 iterations, 6-dof arm) on random
reachable targets, CUDA events around one launch each.  Reference figures for the same calls (probed in the
build container with the reference's own numba code on one core): DiffIK.run 0.88 ms/call, GradIK.run 18.4 ms/call.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from av_aloha_b200 import capi, kinematics, model_io

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = capi.Model(model_io.model_path("slot_insertion", 3), 0)
rng = np.random.default_rng(0)
REF_MS = {"DiffIK": 0.88, "GradIK": 18.4}
for name, arm, ctl in (("DiffIK", "middle", kinematics.DiffIK(model, "middle", **kinematics.DIFFIK_SIM)),
                       ("GradIK", "left", kinematics.GradIK(model, "left", **kinematics.GRADIK_SIM))):
    a = kinematics._arm(arm)
    rngq = model.ik_range(a)
    nd = len(rngq)
    home = np.array([0, -0.8, 0.8, 0, 0.5, 0, 0] if a == 2 else [0, -0.082, 1.06, 0, -0.953, 0])[:nd]
    q = np.clip(home + rng.normal(0, 0.3, (n, nd)), rngq[:, 0], rngq[:, 1]).astype(np.float32)
    tgt = np.clip(q + rng.normal(0, 0.1, (n, nd)), rngq[:, 0], rngq[:, 1]).astype(np.float32)
    T = kinematics.create_fk_fn(model, a)(torch.as_tensor(tgt, device="cuda"))                 # reachable target poses
    pos = T[:, :3, 3].contiguous()
    R = T[:, :3, :3]
    w = torch.sqrt(torch.clamp(1 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2], min=1e-12)) / 2      # rotation -> quaternion (w > 0 branch)
    quat = torch.stack([w, (R[:, 2, 1] - R[:, 1, 2]) / (4 * w), (R[:, 0, 2] - R[:, 2, 0]) / (4 * w),
                        (R[:, 1, 0] - R[:, 0, 1]) / (4 * w)], dim=1).contiguous()
    qd = torch.as_tensor(q, device="cuda")
    for _ in range(2):
        out = ctl.run(qd, pos, quat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ctl.run(qd, pos, quat); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    err = (kinematics.create_fk_fn(model, a)(out)[:, :3, 3] - pos).norm(dim=1)
    print(f"{name} ({arm}, {nd} dof): {n} problems in {ms:.2f} ms -> {n / ms * 1e3:.3e} problems/s "
          f"({ms / n * 1e3:.3f} us each; reference numba on one core: {REF_MS[name]} ms each = {1e3 / REF_MS[name]:.0f}/s); "
          f"median position error after the call {err.median().item() * 1e3:.2f} mm")
