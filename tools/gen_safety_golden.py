"""Generate tests/golden/safety_golden.npz by RUNNING THE REFERENCE'S OWN create_safety_fn / safety
(data_collection_scripts/kinematics.py:54-135, unmodified) behind the same mujoco stub / fake physics as
tools/gen_ik_golden.py.  Cases cover every early return of the reference, in its order.

    python tools/gen_safety_golden.py        # ~1 min (numba JIT)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_ik_golden as g  # noqa: E402  (installs the stub, imports the reference modules)

MESSAGES = ("", "Joint tracking safety margin exceeded", "Joint limit safety margin exceeded",
            "End effector position outside bounds", "End effector action position outside bounds",
            "End effector pose tracking safety margin exceeded")
BOUNDS = {0: [[-0.75, 0.05], [-0.45, 0.45], [-0.05, 0.6]], 1: [[-0.05, 0.75], [-0.45, 0.45], [-0.05, 0.6]],
          2: [[-0.4, 0.4], [-0.9, 0.2], [0.0, 0.9]]}        # synthetic workspace boxes around each arm's reach


def main():
    avm = g.model_io.load_avm(g.model_io.model_path("slot_insertion", 3))
    rng = np.random.default_rng(77)
    out = {}
    for arm in range(3):
        phys, n = g.arm_physics(avm, arm)
        joints = list(range(n))
        fk = g.ref_kin.create_fk_fn(phys, joints, "site")
        fn = g.ref_kin.create_safety_fn(phys, joints, "site", BOUNDS[arm], joint_limit_safety_margin=0.01,
                                        joint_tracking_safety_margin=1.0, eef_pos_tracking_safety_margin=0.2,
                                        eef_rot_tracking_safety_margin=3.0)
        lo, hi = avm["ik_range"][arm, :n, 0], avm["ik_range"][arm, :n, 1]
        N = 96
        q = rng.uniform(lo + 0.02, hi - 0.02, size=(N, n))
        q[: N // 2] = np.clip(g.f32(np.array([0, -0.082, 1.06, 0, -0.953, 0, 0][:n] if arm < 2 else [0, -0.8, 0.8, 0, 0.5, 0, 0]))
                              + rng.normal(0, 0.25, size=(N // 2, n)), lo + 0.02, hi - 0.02)
        ctrl = q + rng.normal(0, 0.05, size=(N, n))
        for i in range(0, N, 8):
            ctrl[i, rng.integers(n)] += rng.choice([-1.0, 1.0]) * 1.3          # joint tracking
        for i in range(1, N, 8):
            j = rng.integers(n)
            q[i, j] = hi[j] - 0.004 if rng.random() < 0.5 else lo[j] + 0.004    # inside the range, inside the margin
            ctrl[i] = q[i]
        q, ctrl = g.f32(q), g.f32(ctrl)
        Ta = np.zeros((N, 4, 4))
        has_T = np.zeros(N, bool)
        code = np.zeros(N, np.int64)
        for i in range(N):
            T = None
            if i % 2 == 0:
                has_T[i] = True
                T = fk(np.clip(q[i] + rng.normal(0, 0.06, n), lo, hi))
                if i % 6 == 0:
                    T = T.copy(); T[:3, 3] += rng.normal(0, 0.25, 3)           # far position / outside the box
                if i % 10 == 0:
                    T = fk(rng.uniform(lo, hi))                               # unrelated pose
                T = np.ascontiguousarray(g.f32(T))
                Ta[i] = T
            ok, msg = fn(q[i].copy(), ctrl[i].copy(), T)
            code[i] = MESSAGES.index(msg)
            assert bool(ok) == (code[i] == 0)
        out[f"q_{arm}"], out[f"ctrl_{arm}"], out[f"T_{arm}"], out[f"hasT_{arm}"], out[f"code_{arm}"] = q, ctrl, Ta, has_T, code
        out[f"bounds_{arm}"] = np.array(BOUNDS[arm], np.float64)
        print(f"arm {arm}: codes", np.bincount(code, minlength=6), flush=True)
    dst = os.path.join(g.ROOT, "tests", "golden", "safety_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
