"""Pins the physics oracle to MuJoCo the day a wheel exists: run this where `import mujoco, dm_control` works (it needs the
reference tree, /root/reference, for the MJCF assets) and commit the files it writes:

    python tools/gen_mujoco_golden.py            ->  tests/golden/mujoco_<task>_<arms>arms.npz   (5 tasks x {2,3} arms)

Per model ~200 states: the reference env is reset (np.random.seed(k)), driven for a random number of steps by a smooth random
joint-target walk with closing grippers (so that fingers, objects and table come into contact), and at the end of it one
`mj_forward` + one `physics.step(nstep=20)` are recorded exactly as reference env.py:203-249 runs them:

    qpos, qvel, ctrl, qacc_warmstart                    the input state (float64)
    qacc, qacc_smooth, qfrc_bias, qfrc_actuator, efc_force-derived qfrc_constraint, M (dense)   after mj_forward
    contact list: ncon, dist, pos, frame[0:3] (normal), geom1, geom2, dim, friction[5], includemargin - dist > 0 (excluded)
    qpos_next, qvel_next                                after physics.step(nstep=20) with the recorded ctrl
    reward                                              env.get_reward() on the stepped state (data.contact after mj_step1)

tests/test_mujoco_golden.py loads whatever files are present (skipped with the reason when there are none) and checks the fp64
oracle against them -- kinematics/inertia exactly, contacts as sets with depth/normal tolerances, qacc to 1e-6 of MuJoCo's own
Newton tolerance, one env.step to 1e-6 -- and the CUDA path against the same vectors on a GPU.  Nothing here can run in the build
container (no MuJoCo wheel in /opt/wheelhouse, no network); the file is committed so that the pin costs one command later.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = os.environ.get("AVSIM_REFERENCE", "/root/reference")
TASKS = {"insert_peg": "InsertPeg", "slot_insertion": "SlotInsertion", "sew_needle": "SewNeedle", "tube_transfer": "TubeTransfer",
         "hook_package": "HookPackage"}


def main(nstates=200):
    import mujoco                                           # noqa: F401  (fails here: that is the point of the probe)
    sys.path.insert(0, os.path.join(REF, "gym_guided_vision"))
    from gym_guided_vision import env as ref_env            # the reference's own env module

    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for task, cls_name in TASKS.items():
        for arms in (2, 3):
            env = getattr(ref_env, cls_name + "Env")(num_arms=arms, cameras=[])
            phys = env.physics
            m, d = phys.model.ptr, phys.data.ptr
            nj = 14 if arms == 2 else 21
            rec = {k: [] for k in ("qpos", "qvel", "ctrl", "warm", "qacc", "qacc_smooth", "qfrc_bias", "qfrc_actuator",
                                   "qfrc_constraint", "M", "ncon", "contacts", "qpos_next", "qvel_next", "reward", "action")}
            rng = np.random.default_rng(1234)
            for k in range(nstates):
                np.random.seed(k)
                obs, _ = env.reset(seed=k)
                a = np.array(obs["agent_pos"][:nj], np.float64)
                target = a + rng.normal(0, 0.25, nj)
                target[[6, 13]] = rng.uniform(0, 1, 2)
                for t in range(int(rng.integers(5, 60))):
                    a += np.clip(target - a, -0.03, 0.03)
                    env.step(a.astype(np.float32))
                mujoco.mj_forward(m, d)
                M = np.zeros((m.nv, m.nv))
                mujoco.mj_fullM(m, M, d.qM)
                con = np.zeros((64, 16))
                for c in range(min(d.ncon, 64)):
                    ct = d.contact[c]
                    con[c, 0] = ct.dist; con[c, 1:4] = ct.pos; con[c, 4:7] = ct.frame[0:3]
                    con[c, 7], con[c, 8], con[c, 9] = ct.geom1, ct.geom2, ct.dim
                    con[c, 10] = float(ct.exclude != 0); con[c, 11:16] = ct.friction
                for key, val in (("qpos", d.qpos), ("qvel", d.qvel), ("ctrl", d.ctrl), ("warm", d.qacc_warmstart), ("qacc", d.qacc),
                                 ("qacc_smooth", d.qacc_smooth), ("qfrc_bias", d.qfrc_bias), ("qfrc_actuator", d.qfrc_actuator),
                                 ("qfrc_constraint", d.qfrc_constraint)):
                    rec[key].append(np.array(val, np.float64))
                rec["M"].append(M); rec["ncon"].append(int(d.ncon)); rec["contacts"].append(con)
                act = a.astype(np.float32)
                _, reward, _, _, _ = env.step(act)
                rec["action"].append(act); rec["reward"].append(int(reward))
                rec["qpos_next"].append(np.array(d.qpos, np.float64)); rec["qvel_next"].append(np.array(d.qvel, np.float64))
            names = {"geom_names": np.array([phys.model.id2name(g, "geom") for g in range(m.ngeom)]),
                     "mujoco_version": np.array(mujoco.__version__)}
            out = os.path.join(ROOT, "tests", "golden", f"mujoco_{task}_{arms}arms.npz")
            np.savez_compressed(out, **{k: np.array(v) for k, v in rec.items()}, **names)
            print("wrote", out)
            env.close()


if __name__ == "__main__":
    try:
        main()
    except ModuleNotFoundError as e:
        raise SystemExit(f"gen_mujoco_golden: {e} -- MuJoCo / dm_control are not importable here (no wheel offline); run this "
                         f"where they are and commit tests/golden/mujoco_*.npz")
