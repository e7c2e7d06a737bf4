"""CPU-only checks of the host side: reset draw order against vectors produced by executing the reference's own reset
code (tools/gen_reset_golden.py), the registry surface, the C-ABI library's exports, and model tables against the
constants the reference states."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reset_golden.json")


@pytest.mark.parametrize("task", ["insert_peg", "slot_insertion", "sew_needle", "tube_transfer", "hook_package"])
def test_reset_draws_follow_reference_rng_order(task):
    from av_aloha_b200 import env, model_io
    gold = json.load(open(GOLD))[task]
    free = model_io.load_names(task, 3)["free_joint"]
    assert sorted(free) == sorted(gold["joints"])
    for case in gold["cases"]:
        np.random.seed(case["seed"])
        for ep in case["episodes"]:                      # consecutive resets share the global stream, dead draws included
            fp = env.reference_reset_draws(task, free)
            for k, j in enumerate(free):
                assert np.array_equal(fp[k], np.array(ep[j][:3])), (task, j)
                assert ep[j][3:] == [1.0, 0.0, 0.0, 0.0]


def test_device_reset_ranges_match_reference_ranges():
    """the Philox device reset samples the same boxes the reference samples (model tables reset_lo / reset_hi)"""
    from av_aloha_b200 import env, model_io
    for task, draws in env._RESET_DRAWS.items():
        avm = model_io.load_avm(model_io.model_path(task, 3))
        free = model_io.load_names(task, 3)["free_joint"]
        table = {j: (lo, hi) for j, lo, hi in draws if j is not None}
        for k, j in enumerate(free):
            lo, hi = table[j]
            if isinstance(lo, str):
                lo, hi = table[lo]
                assert avm["reset_draw"][k] == free.index(table[j][0])      # shares the other joint's draw
            assert np.allclose(avm["reset_lo"][k], lo) and np.allclose(avm["reset_hi"][k], hi)


def test_registry_matches_reference_ids():
    from av_aloha_b200 import env
    ids = [e["id"] for e in env.ENVS]
    assert len(ids) == 10 and len(set(ids)) == 10
    for name in ("InsertPeg", "SlotInsertion", "SewNeedle", "TubeTransfer", "HookPackage"):
        for arms in (2, 3):
            e = next(x for x in env.ENVS if x["id"] == f"gym_guided_vision/{name}-{arms}Arms-v0")
            assert e["kwargs"]["num_arms"] == arms
            assert len(e["kwargs"]["cameras"]) == (6 if arms == 3 else 4)
            assert (e["kwargs"]["observation_height"], e["kwargs"]["observation_width"]) == (480, 640)
    assert env.SIM_PHYSICS_ENV_STEP_RATIO == 20
    assert env.GuidedVisionEnv.metadata["render_fps"] == pytest.approx(25.0)
    # ... and entry by entry against what the reference's own __init__.py registers (recorded by executing it:
    # tools/gen_registry_golden.py): ids, class names, camera LISTS IN ORDER (the 2-arm ids carry no zed cameras)
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "registry_golden.json")))
    assert len(gold) == 10
    for g in gold:
        e = next(x for x in env.ENVS if x["id"] == g["id"])
        assert e["kwargs"] == g["kwargs"], g["id"]
        assert env._TASK_CLASSES[e["task"]].__name__ == g["entry_point"].split(":")[1]
        assert g["nondeterministic"] is True


def test_cabi_exports_every_declared_symbol():
    """libavsim.so loads without a GPU and exports every function include/avsim.h declares (no compute call here)."""
    from av_aloha_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "avsim.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(avsim_[a-z_0-9]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/avsim.h but not exported"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    # error path without a device: null arguments are rejected, message is readable
    lib.avsim_last_error.restype = ctypes.c_char_p
    lib.avsim_model_dim.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    assert lib.avsim_model_dim(None, b"nq") < 0
    assert b"null" in lib.avsim_last_error()


def test_model_tables_match_reference_constants():
    """sizes and actuation constants SURVEY.md 8 / Appendix B derive from aloha_sim.xml"""
    from av_aloha_b200 import model_io
    for task, nq, nv in (("slot_insertion", 37, 35), ("tube_transfer", 44, 41), ("hook_package", 37, 35)):
        for arms in (2, 3):
            m = model_io.load_avm(model_io.model_path(task, arms))
            assert len(m["qpos0"]) == nq and len(m["dof_damping"]) == nv and len(m["act_kp"]) == 21
            assert m["timestep"][0] == 0.002 and m["impratio"][0] == 100 and m["noslip_iterations"][0] == 3
            assert m["num_arms"][0] == arms
    m = model_io.load_avm(model_io.model_path("slot_insertion", 3))
    assert np.allclose(m["act_kp"][:7], [43, 265, 227, 78, 37, 10.4, 2000])
    assert np.allclose(m["act_ctrl_lo"][6], 0.002) and np.allclose(m["act_ctrl_hi"][6], 0.037)
    assert np.allclose(sorted(set(np.round(m["dof_frictionloss"][m["dof_frictionloss"] > 0], 3))), [1.15, 2.0])
    assert (m["dof_frictionloss"] > 0).sum() == 6 and len(m["eq_dof1"]) == 2
    # home pose of the middle arm parks out of view in the 2-arm model (reference env.py:60-62)
    m2 = model_io.load_avm(model_io.model_path("slot_insertion", 2))
    names = model_io.load_names("slot_insertion", 2)
    b = names["body"].index("middle_base_link")
    assert np.allclose(m2["body_pos"][b], [0, -2.4, -0.4])


def test_oracle_is_thread_safe(slot_model_path):
    """bench.py's CPU legs step oracle environments from a thread pool (ctypes releases the GIL): concurrent stepping must
    give bit-identical results to serial stepping (the oracle once kept scratch arrays in function-level statics)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import HOME, OracleEnv, OracleModel

    om = OracleModel(slot_model_path)
    act = HOME.copy()
    act[6] = act[13] = 1.0

    def run(e):
        o = OracleEnv(om)
        o.set_solver("pgs"); o.set_options(max_iter=10, tol=0.0)
        o.reset(free_pos=np.array([[0.01 * e, 0.12, 0.0], [0.02, -0.05 + 0.005 * e, 0.0]]))
        a = act.copy()
        a[0] += 0.05 * e
        for _ in range(3):
            o.step(a)
        return o.qpos.copy()

    serial = [run(e) for e in range(6)]
    with ThreadPoolExecutor(max_workers=6) as ex:
        threaded = list(ex.map(run, range(6)))
    for a, b in zip(serial, threaded):
        assert np.array_equal(a, b)


def test_ik_zero_pose_geometry_matches_an_independent_walk_of_the_mjcf():
    """ik_w0 / ik_p0 / ik_site0 of the compiled model -- what the reference's create_fk_fn reads from mujoco at q = 0
    (kinematics.py:9-15) and what the IK goldens were generated from -- against tests/golden/ik_geometry.json, computed from
    aloha_sim.xml by a separate XML walk (tools/gen_ik_geometry_golden.py: xml.etree + scipy, no code shared with the compiler)."""
    import json
    from av_aloha_b200 import model_io
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ik_geometry.json")))
    for task, arms in (("slot_insertion", 3), ("hook_package", 3), ("insert_peg", 3)):
        avm = model_io.load_avm(model_io.model_path(task, arms))
        for a, arm in enumerate(("left", "right", "middle")):
            n = len(gold[arm]["joints"])
            assert int(avm["ik_ndof"][a]) == n == (6, 6, 7)[a]
            assert np.abs(np.array(gold[arm]["w0"]) - avm["ik_w0"][a, :n]).max() <= 1e-12
            assert np.abs(np.array(gold[arm]["p0"]) - avm["ik_p0"][a, :n]).max() <= 1e-12
            assert np.abs(np.array(gold[arm]["site0"]) - avm["ik_site0"][a]).max() <= 1e-12
    w = np.array(gold["left"]["w0"])
    assert np.allclose(np.linalg.norm(w, axis=1), 1.0) and np.allclose(np.abs(w[0]), [0, 0, 1])      # the waist turns about z


def test_link_inertials_match_an_independent_walk_of_the_mjcf():
    """mass, centre of mass and body-frame inertia tensor of the 28 robot links with an explicit <inertial> (aloha_sim.xml:121 ff.;
    diaginertia rotated by the inertial quat) against the same independent XML walk"""
    import json
    from av_aloha_b200 import model_io
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ik_geometry.json")))["inertials"]
    assert len(gold) == 28
    for task, arms in (("slot_insertion", 3), ("tube_transfer", 3), ("slot_insertion", 2)):
        avm, names = model_io.load_avm(model_io.model_path(task, arms)), model_io.load_names(task, arms)["body"]
        for name, d in gold.items():
            b = names.index(name)
            assert abs(avm["body_mass"][b] - d["mass"]) <= 1e-12 and np.abs(avm["body_ipos"][b] - d["ipos"]).max() <= 1e-12
            assert np.abs(avm["body_inertia"][b] - d["inertia"]).max() <= 1e-12 * max(1.0, np.abs(d["inertia"]).max())


def test_compiled_hulls_match_the_stl_support_functions():
    """Every collision hull against support values computed straight from the reference's STL vertices by an independent reader
    (tools/gen_hull_golden.py -> tests/golden/hull_support.json; 48 Fibonacci directions).  Hull vertices are stored relative to
    the hull centroid, so widths h(d) + h(-d) are compared: exact hulls to 2e-9 m (float32 STL coordinates), the thinned /
    merged fine hulls from inside by at most 0.15 mm.  The absolute placement is checked through
    the compiled geom pose of the finger geoms, whose XML pose (aloha_sim.xml:179, 192) is restated here."""
    import json
    from av_aloha_b200 import model_io
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "hull_support.json")))
    n = 48
    k = np.arange(n) + 0.5
    phi, z = np.pi * (1 + 5 ** 0.5) * k, 1 - 2 * k / n
    D = np.stack([np.sqrt(1 - z * z) * np.cos(phi), np.sqrt(1 - z * z) * np.sin(phi), z], axis=1)
    avm, names = model_io.load_avm(model_io.model_path("slot_insertion", 3)), model_io.load_names("slot_insertion", 3)
    assert set(names["hull"]) == set(gold)
    thinned = 0
    for h, name in enumerate(names["hull"]):
        v = avm["hull_vert"][avm["hull_adr"][h]:avm["hull_adr"][h] + avm["hull_num"][h]]
        proj = v @ D.T
        err = (proj.max(axis=0) - proj.min(axis=0)) - (np.array(gold[name]["support"]) + np.array(gold[name]["support_neg"]))
        assert err.max() <= 2e-9 and err.min() >= -1.5e-4, (name, err.min(), err.max())      # never outside the true hull
        thinned += err.min() < -2e-9
        assert len(v) <= 642
    assert thinned <= 6

    def quat_mat(q):                                          # (w, x, y, z), normalised as MuJoCo does
        w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                         [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                         [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])

    for geom, pos, quat in (("left_left_finger", [0.0141637, 0.0211727, 0.06], [1, 1, 1, -1]),
                            ("left_right_finger", [0.0141637, -0.0211727, 0.0597067], [1, -1, -1, -1])):
        gi = names["geom"].index(geom)
        h = int(avm["geom_hull"][gi])
        v = avm["hull_vert"][avm["hull_adr"][h]:avm["hull_adr"][h] + avm["hull_num"][h]]
        R, Rc = quat_mat(quat), quat_mat(avm["geom_quat"][gi])
        Dw = D @ R.T                                          # mesh-frame directions seen from the body frame
        got = Dw @ avm["geom_pos"][gi] + ((v @ Rc.T) @ Dw.T).max(axis=0)
        want = Dw @ np.array(pos) + np.array(gold[names["hull"][h]]["support"])
        assert np.abs(got - want).max() <= 2e-9, geom
